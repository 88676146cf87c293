"""Candidate slack (k_keep - k) against flagged rows and time (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from palantir_b200 import _device as dv
from palantir_b200.datasets import synth_branching
n = 1_000_000
x, _ = synth_branching(n, 100)
xd = dv.to_device(x)
prep = dv.knn_prepare(xd)
ref = None
for slack in (12, 8, 6, 4):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx, dist, stats = dv.knn_expanded(prep, 31, slack=slack)
        e1.record(); torch.cuda.synchronize()
    same = "" if ref is None else f" identical to slack 12: {bool(torch.equal(idx, ref[0]) and torch.equal(dist, ref[1]))}"
    if ref is None: ref = (idx.clone(), dist.clone())
    print(f"slack {slack}: {e0.elapsed_time(e1):.1f} ms  {stats}{same}", flush=True)
