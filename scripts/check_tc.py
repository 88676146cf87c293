"""Validate the tcgen05 kNN engine against the SIMT engine and FP64 (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv
from palantir_b200.datasets import synth_branching
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
x, _ = synth_branching(n, d)
xd = dv.to_device(x)
res = {}
for eng in ("simt", "tc"):
    dv.KNN_ENGINE = eng
    prep = dv.knn_prepare(xd)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        idx, dist, st = dv.knn_expanded(prep, 31)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(eng, f"{dt*1e3:.1f} ms", st, flush=True)
    res[eng] = (idx.cpu().numpy(), dist.cpu().numpy())
same = sum(set(a) == set(b) for a, b in zip(res["simt"][0], res["tc"][0]))
print("rows with equal neighbour sets:", same, "/", n, " max |ddist|", np.abs(np.sort(res["simt"][1], 1) - np.sort(res["tc"][1], 1)).max())

# --- accuracy of the tensor-core scores against FP64 on the same centred FP32 points ---
dv.KNN_ENGINE = "tc"
prep = dv.knn_prepare(xd)
nq = min(n, 4096); kk = 43
cand = dv.empty((nq, kk), dv.I32); score = dv.empty((nq, kk), dv.F32); thr = dv.empty((nq,), dv.F32)
lists = dv.empty((dv.round_up(nq, 256) * 64,), dv.I64); vis = dv.zeros((1,), dv.I64)
dv.call("pb200_knn_candidates_tc", prep.xtc, prep.ld, prep.yn, prep.xcn, prep.box_lo, prep.box_hi, 0, nq, prep.d, kk,
        lists, cand, score, thr, vis)
torch.cuda.synchronize()
perm = prep.perm.cpu().numpy() if prep.perm is not None else np.arange(n)
mean = prep.mean.cpu().numpy()
xc = (x.astype(np.float64) - mean).astype(np.float32).astype(np.float64)[perm]       # work order, FP32-rounded centred
c = cand.cpu().numpy(); s = score.cpu().numpy().astype(np.float64)
ok = c >= 0
rows = np.repeat(np.arange(nq), kk).reshape(nq, kk)
true = (xc[c.clip(0)] ** 2).sum(-1) - 2.0 * np.einsum("ijk,ik->ij", xc[c.clip(0)], xc[:nq])
err = np.abs(true - s)[ok]
nx = np.sqrt((xc[:nq] ** 2).sum(1)); R = np.sqrt((xc ** 2).sum(1).max())
u = 2.0 ** -24
eps = 1.05 * u * (2.0 * (16 + 4 * 3 * ((d + 7) // 8)) * nx * R + 4 * R * R)
print("TC score error: max", err.max(), "mean", err.mean(), " bound min", eps.min(), " max err/bound", (np.abs(true - s) / eps[:, None])[ok].max())
