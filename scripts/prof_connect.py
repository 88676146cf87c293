"""Is the reconnect loop (core.py:806-842) taken on the benchmark graph, and what does it cost? (dev tool)"""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import pipeline, _device as dv, _device_core as dc
from palantir_b200.datasets import synth_branching
n = 1_000_000
x, info = synth_branching(n, 100)
out = pipeline.run_device(dv.to_device(x), info["early"], list(info["terminals"].values()))
data_t = dc.minmax_scale(out["multiscale"])
perm, inv, work = dc.morton_order(data_t)
idx, dist, st = dc.knn_table_boxed(work, 30)
print("tiles scanned per CTA", int(st["scanned"].item()) / (n / 128))
g = dc.undirected_graph(idx, dist)
start = int(inv[int(out["waypoints"][0])].item())
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nu, label = dc.unreachable_count(g, start)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"unreachable from start: {nu}  ({1e3*(t1-t0):.1f} ms)")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g2 = dc.connect_graph(idx, dist, work, start)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"connect_graph {1e3*(t1-t0):.1f} ms, warnings {len(w)}, arcs {g.nnz} -> {g2.nnz}")
src = torch.tensor([start], dtype=torch.int64, device=idx.device)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    D, _ = dc.sssp(g2, src)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"single-source sssp {1e3*(t1-t0):.1f} ms")
