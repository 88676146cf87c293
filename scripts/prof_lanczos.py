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
order = st.pop("perm", None)
kern = dv.adaptive_kernel(idx, dist, 30, 0.0)
kern.order = order
tvals, svals, dis, dsq = dv.transition_and_symmetric(kern)
np.random.seed(0)
v0 = np.random.rand(n)
start = dv.empty((n,), dv.F64)
dv.call("pb200_vec_mul", dv.to_device(v0), dsq, start, n)
for batch in (20, 20):
    torch.cuda.synchronize(); t0 = time.time()
    lam, U, inf = dv.lanczos_topk(kern, svals, dis, start, 10, batch=batch)
    torch.cuda.synchronize()
    print(f"batch {batch}: {1e3 * (time.time() - t0):.1f} ms, steps {inf['steps']} (windowed {dv.LANCZOS_WINDOWED}, fp32 shadow {dv.LANCZOS_FP32_SHADOW})")
    for m, r in inf["history"]:
        print(f"   m={m:4d} max relative residual {r:.3e}")
    print("eigenvalues", lam)

# SpMV alone: plain kernel on the input-order matrix vs windowed kernel on the relabelled one
perm = order.long(); inv = torch.empty_like(order); inv[perm] = torch.arange(n, dtype=order.dtype, device=order.device)
lens = (kern.indptr[1:] - kern.indptr[:-1])[perm]
ip = torch.zeros((n + 1,), dtype=torch.int32, device=order.device); ip[1:] = torch.cumsum(lens, 0)
ind_p = dv.empty((kern.nnz,), dv.I32); val_p = dv.empty((kern.nnz,), dv.F64)
dv.call("pb200_csr_permute", kern.indptr, kern.indices, svals, n, order, inv, ip, ind_p, val_p)
xv = torch.rand(n, dtype=torch.float64, device=order.device)
y0 = dv.empty((n,), dv.F64); y1 = dv.empty((n,), dv.F64); y2 = dv.empty((n,), dv.F64)
xp = xv[perm].contiguous()
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_plain = timeit(lambda: dv.call("pb200_csr_spmv_f64", kern.indptr, kern.indices, svals, n, kern.nnz, xv, y0))
t_plain_p = timeit(lambda: dv.call("pb200_csr_spmv_f64", ip, ind_p, val_p, n, kern.nnz, xp, y1))
t_win = timeit(lambda: dv.call("pb200_csr_spmv_window_f64", ip, ind_p, val_p, n, kern.nnz, xp, y2))
alg = 12 * kern.nnz + 4 * (n + 1) + 16 * n
print(f"SpMV plain (input order) {t_plain:.3f} ms = {alg / t_plain / 1e6:.0f} GB/s; plain (Morton labels) {t_plain_p:.3f} ms = {alg / t_plain_p / 1e6:.0f} GB/s; "
      f"windowed (Morton labels) {t_win:.3f} ms = {alg / t_win / 1e6:.0f} GB/s")
print("max |windowed - plain(Morton)|", float((y2 - y1).abs().max()), " vs input order (permuted back)", float((y2 - y0[perm]).abs().max()))
