"""Does a Morton relabelling of the cells speed up the SSSP stage? (experiment)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv, _device_core as dc, _native, pipeline
from palantir_b200.datasets import synth_branching
n = 1_000_000
x, info = synth_branching(n, 100)
out = pipeline.run_device(dv.to_device(x), info["early"], list(info["terminals"].values()))
data_t = dc.minmax_scale(out["multiscale"])
perm, inv, work = dc.locality_order(data_t)
idx, dist, st = dc.knn_table_sorted(work, 30)
wp_sorted = inv.cpu().numpy().astype(np.int64)[out["waypoints"]]
lib = _native.load()
def run(tag, g, src):
    mean_w = float(g.data.mean().item())
    torch.cuda.synchronize(); t0 = time.time()
    D, stats = dc.sssp(g, src, delta=4 * mean_w, want_stats=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    s = stats.double().mean(0).cpu().numpy()
    print(f"{tag}: {dt*1e3:.1f} ms rounds {s[0]:.0f} relax/arcs {s[1]/g.nnz:.2f}", flush=True)
    return D
g1 = dc.undirected_graph(idx, dist)
D1 = run("sorted by column 0", g1, dv.to_device(wp_sorted))
# Morton order over the first three columns of the sorted data
m = min(3, work.shape[1])
proj = work[:, :m].t().contiguous()
if m < 3:
    proj = torch.cat([proj, torch.zeros((3 - m, n), dtype=proj.dtype, device=proj.device)])
lo = proj.amin(1).cpu(); hi = proj.amax(1).cpu(); ext = float((hi - lo).max())
nbytes = int(lib.pb200_argsort_scratch_bytes(n)); scratch = dv.empty((nbytes,), torch.uint8)
perm2, inv2 = dv.empty((n,), dv.I32), dv.empty((n,), dv.I32)
dv.call("pb200_morton_order", proj, n, float(lo[0]), float(lo[1]), float(lo[2]), (2**21 - 1) / ext, perm2, inv2, scratch, nbytes)
p2 = perm2.long(); i2 = inv2.long()
idx2 = i2[idx.long()[p2]].int().contiguous()
dist2 = dist[p2].contiguous()
g2 = dc.undirected_graph(idx2, dist2)
wp2 = i2[torch.from_numpy(wp_sorted).to(i2.device)].contiguous()
D2 = run("Morton order (3 columns)", g2, wp2)
print("same distances:", bool(torch.equal(D2[:, i2], D1)) if D1.shape[0] < 2000 else "skipped")
