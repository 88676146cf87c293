"""Step-to-step variation of the stage times (dev tool)."""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import pipeline, _device as dv
from palantir_b200.datasets import synth_branching
n = 1_000_000
x, info = synth_branching(n, 100)
xd = dv.to_device(x)
def smi():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception as e:
        return str(e)
for i in range(9):
    dv.PROFILE = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = pipeline.run_device(xd, info["early"], list(info["terminals"].values()))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    dv.flush_profile()
    p = dv.PROFILE; dv.PROFILE = None
    keys = ["knn_candidates", "knn_rerank", "kernel_build", "lanczos_steps", "knn_direct", "sssp", "refine", "absorption_solve", "project"]
    print(f"step {i}: {1e3*dt:7.1f} ms  " + " ".join(f"{k}={p.get(k, 0):.0f}" for k in keys) + f" sum={sum(p.get(k,0) for k in keys):.0f} | {smi()}", flush=True)
    del out
