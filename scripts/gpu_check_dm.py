"""Quick GPU sanity run of the diffusion-map stages against the golden fixtures (dev tool)."""
import sys, time, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np, torch
import palantir_b200 as pb
from palantir_b200 import _device as dv, utils
from palantir_b200.datasets import synth_branching

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def check(name):
    z = np.load(os.path.join(G, name))
    x = z["x"]; knn = int(z["knn"]); n = x.shape[0]
    xd = dv.to_device(x)
    idx, dist, stats = dv.knn_like_sklearn(xd, min(knn + 1, n))
    idx = idx.cpu().numpy(); dist = dist.cpu().numpy()
    ref_i, ref_d = z["knn_ind"], z["knn_dist"]
    same_rows = sum(set(a) == set(b) for a, b in zip(idx, ref_i))
    print(name, "knn", stats, "rows with equal sets", same_rows, "/", n,
          "max |ddist|", np.abs(np.sort(dist, 1) - np.sort(ref_d, 1))[:, 1:].max(), flush=True)
    import pandas as pd
    res = utils.run_diffusion_maps(pd.DataFrame(x), knn=knn, kernel_backend="sklearn")
    K = res["kernel"]
    print("  K indptr eq", np.array_equal(K.indptr, z["k_indptr"]), "indices eq", np.array_equal(K.indices, z["k_indices"]),
          "max rel dK", (np.abs(K.data - z["k_data"]) / z["k_data"]).max() if K.nnz == z["k_data"].size else "nnz differ")
    if K.nnz == z["k_data"].size:
        print("  max rel dT", (np.abs(res["T"].data - z["t_data"]) / z["t_data"]).max())
    ev = res["EigenValues"].values
    print("  evals rel err vs tight", np.abs(ev - z["evals_tight"]).max(), "vs asis", np.abs(ev - z["evals_asis"]).max(), utils.LAST_INFO)
    from scipy.linalg import subspace_angles
    V = res["EigenVectors"].values
    print("  subspace angle vs tight", subspace_angles(V, z["evecs_tight"]).max(), "asis-vs-tight", subspace_angles(z["evecs_asis"], z["evecs_tight"]).max())
    cos = np.abs(np.sum(V * z["evecs_tight"], 0)); print("  per-vector |cos| min", cos.min(), flush=True)

print(torch.cuda.get_device_name(0))
for f in ["dm_f32_n1500_d20.npz", "dm_f64_n800_d24.npz", "dm_f64_n600_d8.npz"]:
    check(f)

# FP32 peak probe
scr = dv.empty((16,), dv.F32)
for _ in range(2):
    dv.call("pb200_fma_peak_f32", scr, 148 * 8, 4096)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dv.call("pb200_fma_peak_f32", scr, 148 * 8, 16384); e1.record(); torch.cuda.synchronize()
fl = 148 * 8 * 1024 * 16384 * 16 * 2
print("FP32 FMA peak TFLOP/s", fl / e0.elapsed_time(e1) / 1e9)

for n, d in [(100_000, 50), (1_000_000, 100)] if "--big" in sys.argv else [(100_000, 50)]:
    x, info = synth_branching(n, d)
    import pandas as pd
    dv.PROFILE = {}
    t = time.time()
    res = utils.run_diffusion_maps(pd.DataFrame(x), knn=30, kernel_backend="sklearn")
    torch.cuda.synchronize(); wall = time.time() - t
    dv.flush_profile()
    print(n, d, "run_diffusion_maps wall", wall, "stages ms", {k: round(v, 2) for k, v in dv.PROFILE.items()}, utils.LAST_INFO, "nnz/row", res["kernel"].nnz / n)
    print("  evals", res["EigenValues"].values)
    if "knn_candidates" in dv.PROFILE:
        print("  kNN TFLOP/s", 2.0 * n * n * d / dv.PROFILE["knn_candidates"] / 1e9)
