"""Wall time of the device-resident pipeline vs the public API, phase by phase (dev tool)."""
import sys, os, io, contextlib, warnings, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np, pandas as pd, torch
from palantir_b200 import utils, core, pipeline, _device as dv
from palantir_b200.datasets import synth_branching, synth_names
n = 1_000_000
x, info = synth_branching(n, 100)
names = synth_names(n)
df = pd.DataFrame(x, index=names)
term = {names[v]: k for k, v in info["terminals"].items()}
xd = dv.to_device(x)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
def dev_step():
    return pipeline.run_device(xd, info["early"], list(info["terminals"].values()))
def api_step(tm):
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = sync(); res = utils.run_diffusion_maps(df, n_components=10, knn=30, kernel_backend="sklearn")
        t1 = sync(); ms = utils.determine_multiscale_space(res)
        t2 = sync(); pr = core.run_palantir(ms, names[info["early"]], terminal_states=term, num_waypoints=1200)
        t3 = sync()
    tm.append((t1 - t0, t2 - t1, t3 - t2))
    return res, ms, pr
for _ in range(2): dev_step()
for _ in range(3):
    t0 = sync(); out = dev_step(); t1 = sync(); print(f"pipeline.run_device wall {1e3*(t1-t0):.1f} ms")
tm = []
r = None
for _ in range(2): r = None; r = api_step(tm)
tm = []
for _ in range(3):
    r = None
    t0 = sync(); r = api_step(tm); t1 = sync()
    print(f"API wall {1e3*(t1-t0):.1f} ms: run_diffusion_maps {1e3*tm[-1][0]:.1f}, determine_multiscale_space {1e3*tm[-1][1]:.1f}, run_palantir {1e3*tm[-1][2]:.1f}")
# the same with a page-locked input matrix
xp = torch.from_numpy(x).pin_memory().numpy()
df = pd.DataFrame(xp, index=names, copy=False)
for _ in range(3):
    r = None
    t0 = sync(); r = api_step(tm); t1 = sync()
    print(f"API wall (pinned input) {1e3*(t1-t0):.1f} ms: run_diffusion_maps {1e3*tm[-1][0]:.1f}, determine_multiscale_space {1e3*tm[-1][1]:.1f}, run_palantir {1e3*tm[-1][2]:.1f}")
