"""Quick GPU sanity run of run_palantir against the golden fixture + timing at scale (dev tool)."""
import sys, time, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np, pandas as pd, torch
import palantir_b200 as pb
from palantir_b200 import _device as dv, utils, core
from palantir_b200.datasets import synth_branching, synth_names
from scipy.stats import spearmanr

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
z = np.load(os.path.join(G, "palantir_n1500_wp200.npz"))
n = z["ms"].shape[0]
names = synth_names(n)
ms = pd.DataFrame(z["ms"], index=names)
term = list(names[z["terminals"]])
pr = core.run_palantir(ms, names[int(z["early"])], terminal_states=term, num_waypoints=int(z["num_waypoints"]), knn=int(z["knn"]))
print("waypoints identical", list(names.get_indexer(pr.waypoints)) == list(z["waypoints"]))
print("pseudotime max abs", np.abs(pr.pseudotime.values - z["pseudotime"]).max(), "spearman", spearmanr(pr.pseudotime.values, z["pseudotime"])[0])
print("fates max abs", np.abs(pr.branch_probs.values - z["fate_probs"]).max(), "cols", list(names.get_indexer(pr.branch_probs.columns)) == list(z["fate_columns"]))
print("entropy max abs", np.abs(pr.entropy.values - z["entropy"]).max(), flush=True)

# reference known answers
k = np.load(os.path.join(G, "ref_known_answers.npz"))
wp = core._max_min_sampling(pd.DataFrame(k["mm_data"]), num_waypoints=9, seed=0)
print("maxmin known answer", list(wp) == list(k["mm_waypoints"]))
b = pd.DataFrame(k["pt_data"]); b.index = pd.Index([f"cell_{i}" for i in range(25)])
pt, W = core._compute_pseudotime(b, b.index[0], knn=5, waypoints=b.index[[0, 5, 10, 15]], n_jobs=1, max_iterations=10)
print("pseudotime known answer", np.abs(pt.values - k["pt_pseudotime"]).max(), "W", np.abs(W.values - k["pt_w"]).max(), flush=True)

# auto terminal states + tiny input like the reference's tests
rng = np.random.default_rng(3)
tiny = pd.DataFrame(rng.random((50, 10)), index=[f"cell_{i}" for i in range(50)])
r = core.run_palantir(tiny, "cell_0")
print("tiny auto-terminal ok", r.branch_probs.shape, r.pseudotime.shape)

for nn, d, nw in ([(100_000, 50, 1200), (1_000_000, 100, 1200)] if "--big" in sys.argv else [(100_000, 50, 1200)]):
    x, info = synth_branching(nn, d)
    names = synth_names(nn)
    dv.PROFILE = {}
    t0 = time.time()
    res = utils.run_diffusion_maps(pd.DataFrame(x, index=names), knn=30, kernel_backend="sklearn")
    msd = utils.determine_multiscale_space(res)
    t1 = time.time()
    term = {names[v]: kk for kk, v in info["terminals"].items()}
    pr = core.run_palantir(msd, names[info["early"]], terminal_states=term, num_waypoints=nw)
    torch.cuda.synchronize(); t2 = time.time()
    dv.flush_profile()
    print(nn, d, "dm wall", t1 - t0, "palantir wall", t2 - t1, "cells/s", nn / (t2 - t0))
    print("  stages ms", {kk: round(v, 1) for kk, v in dv.PROFILE.items()})
    print("  ", core.LAST_INFO, "ms dims", msd.shape[1])
    print("  spearman(pt, true time)", spearmanr(pr.pseudotime.values, info["time"])[0], "fates shape", pr.branch_probs.shape,
          "own-terminal fate", [float(pr.branch_probs.loc[c].max()) for c in term])
