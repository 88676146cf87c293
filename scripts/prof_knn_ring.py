"""tcgen05 kNN candidate kernel: ring depth against candidate-buffer size (they share the 227 KB of shared memory),
at benchmark size (dev tool).  PB200_TC_NST / PB200_TC_CAP are read by the library at every launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv
from palantir_b200.datasets import synth_branching
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x, _ = synth_branching(n, 100)
xd = dv.to_device(x)
prep = dv.knn_prepare(xd)
if os.environ.get("PB200_TC_DEBUG") == "4":
    # half of the columns processed: wrong candidates, but thresholds still tighten and tiles are still pruned
    kk = 43
    cand = dv.empty((n, kk), dv.I32); score = dv.empty((n, kk), dv.F32); thr = dv.empty((n,), dv.F32)
    lists = dv.empty((256,), dv.I64); vis = dv.zeros((1,), dv.I64)
    for rep in range(3):
        vis.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dv.call("pb200_knn_candidates_tc", prep.xtc, prep.ld, prep.yn, prep.xcn, prep.box_lo, prep.box_hi, 0, n,
                prep.d, kk, lists, cand, score, thr, vis)
        e1.record(); torch.cuda.synchronize()
        tiles = float(vis.item())
        print(f"debug 4: candidate kernel {e0.elapsed_time(e1):.1f} ms, {tiles:.3e} tiles, "
              f"{3 * 2 * 128 * 128 * 104 * tiles / (e0.elapsed_time(e1) * 1e-3) / 1e12:.0f} TFLOP/s executed", flush=True)
    sys.exit(0)
if os.environ.get("PB200_TC_DEBUG", "0") != "0":
    # a debug mode of the kernel (1 = accumulators released unread, 2 = no candidate pushes): candidate kernel alone
    kk = 43
    cand = dv.empty((n, kk), dv.I32); score = dv.empty((n, kk), dv.F32); thr = dv.empty((n,), dv.F32)
    lists = dv.empty((256,), dv.I64); vis = dv.zeros((1,), dv.I64)
    rings = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [5, 2, 3, 4, 6, 8]
    for nst, cap in [(r, 128 if r <= 5 else 80) for r in rings]:
        os.environ["PB200_TC_NST"], os.environ["PB200_TC_CAP"] = str(nst), str(cap)
        ts = []
        for rep in range(3):
            vis.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dv.call("pb200_knn_candidates_tc", prep.xtc, prep.ld, prep.yn, prep.xcn, prep.box_lo, prep.box_hi, 0, n,
                    prep.d, kk, lists, cand, score, thr, vis)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        tiles = float(vis.item())
        tf = 3 * 2 * 128 * 128 * 104 * tiles / (min(ts) * 1e-3) / 1e12
        print(f"debug {os.environ['PB200_TC_DEBUG']}: ring {nst}: candidate kernel {min(ts):.1f} ms, {tiles:.3e} tiles, {tf:.0f} TFLOP/s executed", flush=True)
    sys.exit(0)
ref = None
configs = [(5, 128, 8), (5, 128, 0), (5, 128, 4), (5, 128, 12), (5, 128, 16), (6, 112, 12), (5, 128, 8)]
if len(sys.argv) > 2:
    configs = [tuple(int(v) for v in c.split(":")) for c in sys.argv[2].split(",")]
for cfg in configs:
    nst, cap, slack = cfg[:3]
    idle = cfg[3] if len(cfg) > 3 else 20
    os.environ["PB200_TC_NST"], os.environ["PB200_TC_CAP"], os.environ["PB200_TC_SLACK"] = str(nst), str(cap), str(slack)
    os.environ["PB200_TC_IDLE"] = str(idle)
    ts = []
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx, dist, st = dv.knn_expanded(prep, 31)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    same = True if ref is None else bool(torch.equal(idx, ref[0]) and torch.equal(dist, ref[1]))
    if ref is None:
        ref = (idx.clone(), dist.clone())
    print(f"ring {nst} x 16 KB, buffers {cap}, slack {slack}, idle margin {idle}: knn_expanded {min(ts):.1f} ms (runs {[round(t, 1) for t in ts]}); "
          f"tiles visited {st.get('tiles_visited_frac')}, flagged rows {st.get('flagged')}; same result {same}", flush=True)
