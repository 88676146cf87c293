"""Config 4 at single-GPU scale: 4 M cells x 50 PCs, one GPU's share of the 5000 waypoints (625), device-resident."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.stats import spearmanr
from palantir_b200 import pipeline, _device as dv
from palantir_b200.datasets import synth_branching
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
nwp = int(sys.argv[2]) if len(sys.argv) > 2 else 625
t0 = time.time(); x, info = synth_branching(n, 50); print(f"data {time.time()-t0:.1f} s", flush=True)
xd = dv.to_device(x)
for rep in range(2):
    dv.PROFILE = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = pipeline.run_device(xd, info["early"], list(info["terminals"].values()), num_waypoints=nwp)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    dv.flush_profile()
    print(f"rep {rep}: {dt:.2f} s = {n/dt:.0f} cells/s; stages ms " + str({k: round(v) for k, v in dv.PROFILE.items() if v >= 5}), flush=True)
    dv.PROFILE = None
pt = out["pseudotime"].cpu().numpy(); fp = out["fate_probs"].cpu().numpy()
print("peak memory GiB", torch.cuda.max_memory_allocated() / 2**30, "knn", out["knn"], "lanczos", out["eigenvalues"][:4])
print("pt range", pt.min(), pt.max(), "spearman vs latent time", spearmanr(pt[::7], info["time"][::7])[0])
print("fates: rows sum min/max", fp.sum(1).min(), fp.sum(1).max(), "terminal rows", [float(fp[p].max()) for p in info["terminals"].values()])
