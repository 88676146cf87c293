"""Do the slow steps coincide with cudaMalloc / cudaFree traffic of the caching allocator? (dev tool)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import pipeline, _device as dv
from palantir_b200.datasets import synth_branching
n = 1_000_000
x, info = synth_branching(n, 100)
xd = dv.to_device(x)
def stats():
    s = torch.cuda.memory_stats()
    return s["num_device_alloc"], s["num_device_free"], s["num_alloc_retries"], s["reserved_bytes.all.current"] / 2**30
prev = stats()
for i in range(7):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = pipeline.run_device(xd, info["early"], list(info["terminals"].values()))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    cur = stats()
    print(f"step {i}: {1e3*dt:7.1f} ms  cudaMalloc +{cur[0]-prev[0]} cudaFree +{cur[1]-prev[1]} retries +{cur[2]-prev[2]} reserved {cur[3]:.1f} GiB")
    prev = cur
    del out
