"""Balanced kNN shards (dev tool): the visit counts of sampled query tiles run to the end against the estimate after
a few tiles, the cuts both give for 2 / 4 / 8 ranks, and what the sampled launch costs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv
from palantir_b200.datasets import synth_branching
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x, _ = synth_branching(n, 100)
prep = dv.knn_prepare(dv.to_device(x))
k_keep = 43
n_qt = (n + 127) // 128
stride = max(1, -(-n_qt // 148))
sample = torch.arange(stride // 2, n_qt, stride, dtype=dv.I32, device="cuda")
ns = int(sample.numel())
cand = dv.empty((ns * 128, k_keep), dv.I32); score = dv.empty((ns * 128, k_keep), dv.F32); thr = dv.empty((ns * 128,), dv.F32)
res = {}
for limit in (0, 48, 24, 12, 6):
    counts = dv.zeros((ns,), dv.I32)
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dv.call("pb200_knn_tc_visit_counts", prep.xtc, prep.ld, prep.yn, prep.xcn, prep.box_lo, prep.box_hi, n, prep.d,
                k_keep, sample, ns, limit, cand, score, thr, counts)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    c = counts.cpu().numpy().astype(np.float64)
    res[limit] = c
    work = np.interp(np.arange(n_qt), sample.cpu().numpy().astype(np.float64), np.maximum(c, 1.0))
    true_work = np.interp(np.arange(n_qt), sample.cpu().numpy().astype(np.float64), np.maximum(res[0], 1.0))
    line = f"limit {limit:3d}: {min(ts):.3f} ms; counts mean {c.mean():.0f} min {c.min():.0f} max {c.max():.0f}; corr with full {np.corrcoef(c, res[0])[0, 1]:.4f};"
    for parts in (2, 4, 8):
        cuts = dv.cuts_at_equal_work(work, parts, n)
        share = [true_work[a // 128:(b + 127) // 128].sum() for a, b in zip(cuts, cuts[1:])]
        line += f" {parts} ranks: slowest/mean {max(share) / (true_work.sum() / parts):.3f}"
    print(line, flush=True)
even = lambda parts: max(true_work[a // 128:(b + 127) // 128].sum() for a, b in zip(range(0, n, n // parts), range(n // parts, n + 1, n // parts))) / (true_work.sum() / parts)
print("equal row counts: slowest/mean", {p: round(even(p), 3) for p in (2, 4, 8)})
