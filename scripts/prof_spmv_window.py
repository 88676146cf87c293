"""How local are the kernel's columns once the cells are in the kNN's Morton order? (experiment)"""
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
prep = dv.knn_prepare(xd)
perm = prep.perm.long()
inv = torch.empty_like(perm); inv[perm] = torch.arange(n, device=perm.device)
rows = torch.repeat_interleave(torch.arange(n, device=perm.device), (kern.indptr[1:] - kern.indptr[:-1]).long())
cols = kern.indices.long()
def report(tag, r, c):
    dlt = (c - r).abs()
    print(tag, "nnz", dlt.numel(), " |col-row| quantiles 50/90/99/99.9 %:",
          [int(v) for v in torch.quantile(dlt[:: max(1, dlt.numel() // 4_000_000)].double(), torch.tensor([.5, .9, .99, .999], dtype=torch.float64, device=dlt.device)).tolist()])
    for w in (2048, 4096, 8192, 12288, 16384, 24576, 50000):
        print(f"   within +-{w}: {float((dlt <= w).double().mean()):.4f}")
    # block-window measure: rows in blocks of R, window = [blockstart - H, blockstart + R + H)
    for R, H in ((1024, 7680), (2048, 7168), (1024, 11264), (4096, 10240)):
        b0 = (r // R) * R
        inside = ((c >= b0 - H) & (c < b0 + R + H)).double().mean()
        print(f"   rows/block {R} halo {H} (window {R + 2 * H} doubles = {(R + 2 * H) * 8 // 1024} KB): inside {float(inside):.4f}")
report("input order ", rows, cols)
report("Morton order", inv[rows], inv[cols])
# ordering on the first principal projection only
