"""Time pb200_knn_candidates_tc alone (PB200_TC_DEBUG=1/2 isolates the MMA pipeline / the fast epilogue)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
from palantir_b200 import _device as dv, _native
from palantir_b200.datasets import synth_branching
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--d", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-prune", action="store_true")
a = ap.parse_args()
x, _ = synth_branching(a.n, a.d)
xd = dv.to_device(x)
dv.KNN_ENGINE = "tc"
prep = dv.knn_prepare(xd, prune=not a.no_prune)
nq, k_keep = a.n, 43
cap = int(_native.load().pb200_knn_tc_list_cap())
cand = dv.empty((nq, k_keep), dv.I32); score = dv.empty((nq, k_keep), dv.F32); thr = dv.empty((nq,), dv.F32)
lists = dv.empty((dv.round_up(nq, 256) * cap,), dv.I64); visited = dv.zeros((1,), dv.I64)
for r in range(a.reps):
    visited.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dv.call("pb200_knn_candidates_tc", prep.xtc, prep.ld, prep.yn, prep.xcn, prep.box_lo, prep.box_hi, 0, nq,
            prep.d, k_keep, lists, cand, score, thr, visited)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tiles = int(visited.item())
    d8 = (a.d + 7) // 8 * 8
    print(f"debug={os.environ.get('PB200_TC_DEBUG', '0')} rep {r}: {ms:.2f} ms  tiles {tiles}  executed TF32 {3 * 2.0 * tiles * 128 * 128 * d8 / ms / 1e9:.1f} TFLOP/s  ns/tile/SM {ms * 1e6 * 148 / max(tiles, 1):.0f}")
