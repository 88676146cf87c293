"""cProfile of the public-API path at benchmark size (dev tool)."""
import sys, os, cProfile, pstats, io, contextlib, warnings, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np, pandas as pd, torch
from palantir_b200 import utils, core
from palantir_b200.datasets import synth_branching, synth_names
n = 1_000_000
x, info = synth_branching(n, 100)
names = synth_names(n)
df = pd.DataFrame(x, index=names)
term = {names[v]: k for k, v in info["terminals"].items()}
def step():
    with contextlib.redirect_stdout(io.StringIO()):
        res = utils.run_diffusion_maps(df, n_components=10, knn=30, kernel_backend="sklearn")
        ms = utils.determine_multiscale_space(res)
        pr = core.run_palantir(ms, names[info["early"]], terminal_states=term, num_waypoints=1200)
    return res, ms, pr
step(); torch.cuda.synchronize()
t0 = time.time(); step(); torch.cuda.synchronize(); print("wall", time.time() - t0)
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
