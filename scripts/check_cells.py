"""Dev check of the cell-decomposed shortest paths (sssp_cells.cu) against the per-source kernel and scipy."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import pipeline, _device as dv, _device_core as dc
from palantir_b200.datasets import synth_branching

def graph_for(n, d):
    x, info = synth_branching(n, d)
    out = pipeline.run_device(dv.to_device(x), info["early"], list(info["terminals"].values()), num_waypoints=64)
    dev = dc.minmax_scale(out["multiscale"])
    perm, inv, work = dc.morton_order(dev)
    idx, dist, _ = dc.knn_table_boxed(work, 30)
    return dc.connect_graph(idx, dist, work, int(inv[info["early"]].item()))

def timed(f, reps=3):
    for _ in range(1):
        r = f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) / reps * 1e3

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 50
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
target = int(sys.argv[4]) if len(sys.argv) > 4 else None
g = graph_for(n, d)
print(f"graph n={g.n} nnz={g.nnz}", flush=True)
rng = np.random.default_rng(0)
src = dv.to_device(rng.choice(n, S, replace=False).astype(np.int64))
if target:
    dc.CELL_TARGET = target
dc.CELL_UNIT_GROW = bool(int(os.environ.get("CELL_UNIT", "0")))
dc.CELL_LOCAL_DELTA = float(os.environ.get("CELL_LDELTA", "1"))
dc.CELL_ALPHA = float(os.environ.get("CELL_ALPHA", "1"))
res = dc.sssp_cells(g, src)
print("cells info", dc.LAST_CELLS, flush=True)
if os.environ.get("CHECK_CELLS_ONCE"):
    torch.cuda.synchronize(); sys.exit(0)
if res is None:
    print("cell path declined"); sys.exit(0)
D, order, _ = res
ref, _ = dc.sssp_per_source(g, src[:min(S, 8)])
got = D[:min(S, 8)][:, dc.empty((0,), dc.I32).new_zeros(0).long()] if False else None
inv = torch.empty_like(order); inv[order.long()] = torch.arange(n, dtype=order.dtype, device=order.device)
got = D[:min(S, 8)][:, inv.long()]
fin = torch.isfinite(ref)
same_inf = bool((torch.isfinite(got) == fin).all())
rel = ((got[fin] - ref[fin]).abs() / ref[fin].clamp_min(1e-300)).max().item()
print(f"vs per-source kernel: finite pattern equal {same_inf}, max rel diff {rel:.3e}", flush=True)
dv.PROFILE = {}
_, ms_new = timed(lambda: dc.sssp_cells(g, src))
dv.flush_profile(); prof = {k: round(v / 4, 3) for k, v in dv.PROFILE.items()}; dv.PROFILE = None
print(f"cells path: {ms_new:.2f} ms for {S} sources; stages {prof}", flush=True)
st = dc.sssp_cells(g, src, want_stats=True)[2]
print("overlay search: rounds mean %.0f, relaxed arcs mean %.3g, advances mean %.0f" % tuple(st.double().mean(0).tolist()), flush=True)
if S <= 1300 and n <= 1_000_000 and not os.environ.get("CELL_SKIP_OLD"):
    _, ms_old = timed(lambda: dc.sssp_per_source(g, src), reps=1)
    print(f"per-source kernel: {ms_old:.2f} ms", flush=True)
