"""Morton vs Hilbert ordering: tiles visited / times of the two pruned kNNs, SpMV column window (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv, _device_core as dc, pipeline
from palantir_b200.datasets import synth_branching
n = 1_000_000
x, info = synth_branching(n, 100)
xd = dv.to_device(x)
print("curve", os.environ.get("PB200_CURVE", "hilbert"))
dv.PROFILE = {}
out = pipeline.run_device(xd, info["early"], list(info["terminals"].values()))
dv.PROFILE = {}
out = pipeline.run_device(xd, info["early"], list(info["terminals"].values()))
dv.flush_profile()
print({k: round(v, 1) for k, v in dv.PROFILE.items() if k in ("knn_candidates", "knn_direct", "sssp", "knn_rerank", "lanczos_steps")}, out["knn"])
dv.PROFILE = None
kern = out["kernel"]
prep = dv.knn_prepare(xd)
perm = prep.perm.long()
inv = torch.empty_like(perm); inv[perm] = torch.arange(n, device=perm.device)
rows = torch.repeat_interleave(torch.arange(n, device=perm.device), (kern.indptr[1:] - kern.indptr[:-1]).long())
r, c = inv[rows], inv[kern.indices.long()]
dlt = (c - r).abs()
for w in (4096, 8192, 12288, 16384, 24576, 50000):
    print(f"   kernel columns within +-{w}: {float((dlt <= w).double().mean()):.4f}")
