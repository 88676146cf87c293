"""Dev probe: run the device pipeline at a given size and bring the scaled multiscale matrix, the waypoints and the
start cell home (gpurun_out/), so the SSSP stage's graph can be rebuilt and studied on the CPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import pipeline, _device as dv, _device_core as dc
from palantir_b200.datasets import synth_branching
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
x, info = synth_branching(n, d)
out = pipeline.run_device(dv.to_device(x), info["early"], list(info["terminals"].values()))
ms = out["multiscale"]
dev = dc.minmax_scale(ms).cpu().numpy()
print("multiscale", dev.shape, "eigenvalues", out["eigenvalues"])
if dev.nbytes > 50 << 20:
    dev = dev.astype(np.float32)
np.save(f"gpurun_out/ms_{n}.npy", dev)
np.save(f"gpurun_out/wp_{n}.npy", out["waypoints"])
