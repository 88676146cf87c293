"""Run the kNN candidate kernel a few times on synthetic data (target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
from palantir_b200 import _device as dv
from palantir_b200.datasets import synth_branching

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=151552)
ap.add_argument("--d", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-prune", action="store_true")
ap.add_argument("--engine", default="simt")
a = ap.parse_args()
x, _ = synth_branching(a.n, a.d)
xd = dv.to_device(x)
dv.KNN_ENGINE = a.engine
prep = dv.knn_prepare(xd, prune=not a.no_prune)
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    idx, dist, stats = dv.knn_expanded(prep, 31)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {r}: {ms:.2f} ms  {2.0*a.n*a.n*a.d/ms/1e9:.2f} TFLOP/s  {stats}")
