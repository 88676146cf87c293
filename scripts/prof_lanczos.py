"""Residual history of the Lanczos solve on the benchmark kernel (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv
from palantir_b200.datasets import synth_branching
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x, info = synth_branching(n, 100)
xd = dv.to_device(x)
idx, dist, st = dv.knn_like_sklearn(xd, 31)
kern = dv.adaptive_kernel(idx, dist, 30, 0.0)
tvals, svals, dis, dsq = dv.transition_and_symmetric(kern)
np.random.seed(0)
v0 = np.random.rand(n)
start = dv.empty((n,), dv.F64)
dv.call("pb200_vec_mul", dv.to_device(v0), dsq, start, n)
for batch in (10,):
    torch.cuda.synchronize(); t0 = time.time()
    lam, U, inf = dv.lanczos_topk(kern, svals, dis, start, 10, tol=1e-13, batch=batch)
    torch.cuda.synchronize()
    print(f"batch {batch}: {1e3 * (time.time() - t0):.1f} ms, steps {inf['steps']}")
    for m, r in inf["history"]:
        print(f"   m={m:4d} max relative residual {r:.3e}")
    print("eigenvalues", lam)
