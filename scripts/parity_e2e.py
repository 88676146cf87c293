"""P-e2e parity report: the whole pipeline on the GPU vs the CPU oracle run as-is
(ARPACK tol 1e-4, the reference's setting) and tight (tol 1e-12), next to the oracle's own
as-is <-> tight self-noise.  Also times the CPU oracle and the GPU API on the same box.

  python scripts/parity_e2e.py [--cells 20000] [--pcs 50] [--waypoints 1200]
"""
import argparse, contextlib, io, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np, pandas as pd, torch
from scipy.linalg import subspace_angles
from scipy.stats import spearmanr
from oracle import palantir_oracle as po
from palantir_b200 import core, utils
from palantir_b200.datasets import synth_branching, synth_names

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=20000); ap.add_argument("--pcs", type=int, default=50)
ap.add_argument("--waypoints", type=int, default=1200); ap.add_argument("--skip-tight", action="store_true")
a = ap.parse_args()
n, d, nw = a.cells, a.pcs, a.waypoints
x, info = synth_branching(n, d)
names = synth_names(n)
term = list(info["terminals"].values())
quiet = lambda: contextlib.redirect_stdout(io.StringIO())

def cpu(tol):
    t0 = time.perf_counter()
    with quiet():
        dm = po.run_diffusion_maps(x, 10, 30, tol=tol)
        ms = po.multiscale_space(dm["EigenVectors"], dm["EigenValues"])
        pr = po.run_palantir(ms, info["early"], terminal=term, num_waypoints=nw)
    return dm, ms, pr, time.perf_counter() - t0

def gpu():
    df = pd.DataFrame(x, index=names)
    with quiet():
        utils.run_diffusion_maps(df.iloc[:2000], kernel_backend="sklearn")     # warm-up (library load, allocator)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with quiet():
        dm = utils.run_diffusion_maps(df, n_components=10, knn=30, kernel_backend="sklearn")
        ms = utils.determine_multiscale_space(dm)
        pr = core.run_palantir(ms, names[info["early"]], terminal_states=list(names[term]), num_waypoints=nw)
    torch.cuda.synchronize()
    return dm, ms, pr, time.perf_counter() - t0

def cmp(tag, ev_a, V_a, ms_a, wp_a, pt_a, fp_a, ev_b, V_b, ms_b, wp_b, pt_b, fp_b):
    k = min(ms_a.shape[1], ms_b.shape[1])
    print(f"{tag:28s} eig rel {np.max(np.abs(ev_a-ev_b)/np.abs(ev_b)):.2e}  angle[:10] {subspace_angles(V_a, V_b).max():.2e}  "
          f"angle[ms] {subspace_angles(ms_a[:, :k], ms_b[:, :k]).max():.2e}  wp overlap {len(set(wp_a) & set(wp_b))/len(wp_b):.3f}  "
          f"pt max|d| {np.max(np.abs(pt_a-pt_b)):.2e} spearman {spearmanr(pt_a, pt_b)[0]:.7f}  "
          f"fates max|d| {np.max(np.abs(fp_a-fp_b)) if fp_a.shape == fp_b.shape else float('nan'):.2e}")

g_dm, g_ms, g_pr, g_t = gpu()
a_dm, a_ms, a_pr, a_t = cpu(1e-4)
print(f"cells {n} pcs {d} waypoints {nw}: GPU API {g_t:.2f} s, CPU oracle as-is {a_t:.1f} s on {os.cpu_count()} cores "
      f"(x{a_t / g_t:.0f})")
G = (g_dm["EigenValues"].values, g_dm["EigenVectors"].values, g_ms.values, names.get_indexer(g_pr.waypoints),
     g_pr.pseudotime.values, g_pr.branch_probs.values)
A = (a_dm["EigenValues"], a_dm["EigenVectors"], a_ms, a_pr["waypoints"], a_pr["pseudotime"], a_pr["fate_probs"])
cmp("GPU vs oracle as-is", *G, *A)
if not a.skip_tight:
    t_dm, t_ms, t_pr, t_t = cpu(1e-12)
    T = (t_dm["EigenValues"], t_dm["EigenVectors"], t_ms, t_pr["waypoints"], t_pr["pseudotime"], t_pr["fate_probs"])
    cmp("GPU vs oracle tight", *G, *T)
    cmp("oracle as-is vs tight (noise)", *A, *T)
    # stage-isolated: same multiscale matrix (the tight oracle's) on both sides
    with quiet():
        pr = core.run_palantir(pd.DataFrame(t_ms, index=names), names[info["early"]], terminal_states=list(names[term]),
                               num_waypoints=nw)
    print(f"stage-isolated run_palantir (tight oracle's multiscale matrix): waypoints identical "
          f"{list(names.get_indexer(pr.waypoints)) == list(t_pr['waypoints'])}, pt max|d| "
          f"{np.max(np.abs(pr.pseudotime.values - t_pr['pseudotime'])):.2e}, fates max|d| "
          f"{np.max(np.abs(pr.branch_probs.values - t_pr['fate_probs'])):.2e}")
