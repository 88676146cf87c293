"""Per-step stage times of the device-resident pipeline at config 3 (dev tool): is a slow average one slow step?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv, pipeline, core
from palantir_b200.datasets import synth_branching
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
x, info = synth_branching(n, 100)
xd = dv.to_device(x)
keys = ["knn_candidates", "eigensolver", "sssp_partition", "sssp_tables", "sssp", "sssp_expand", "absorption_solve"]
for s in range(steps):
    dv.PROFILE = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pipeline.run_device(xd, info["early"], list(info["terminals"].values()))
    e1.record(); torch.cuda.synchronize()
    dv.flush_profile()
    cells = core.LAST_INFO.get("sssp", {}).get("cells", {})
    print(f"step {s}: {e0.elapsed_time(e1):7.1f} ms  " + "  ".join(f"{k} {dv.PROFILE.get(k, 0.0):6.1f}" for k in keys)
          + f"  boundary {cells.get('boundary')} tables {cells.get('table_doubles')}", flush=True)
    dv.PROFILE = None
    del out
