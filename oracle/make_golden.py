"""Pin the oracle against the reference and write the fixtures under tests/golden/.

Run in the build container (needs /root/reference):  python -m oracle.make_golden

For every case the reference's OWN functions (loaded by ref_shim from
/root/reference/src/palantir) are run on seeded inputs; this script asserts that
oracle/palantir_oracle.py reproduces them and stores the reference outputs as
small .npz fixtures.  tests/ compares both the oracle (CPU) and the CUDA path
(GPU) with these files; /root/reference is never read at test time.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import palantir_oracle as po, ref_shim  # noqa: E402
from palantir_b200.datasets import synth_branching, synth_names  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _csr_parts(m):
    m = m.tocsr()
    m.sort_indices()
    return dict(indptr=m.indptr.astype(np.int32), indices=m.indices.astype(np.int32), data=m.data)


def case_diffusion(core, utils, utils_tight, name, n, d, dtype, knn=30, seed=0):
    from sklearn.neighbors import NearestNeighbors

    x, info = synth_branching(n, d, seed=seed, dtype=dtype)
    names = synth_names(n)
    df = pd.DataFrame(x, index=names)
    res = utils.run_diffusion_maps(df, n_components=10, knn=knn, kernel_backend="sklearn")
    res_t = utils_tight.diffusion_maps_from_kernel(res["kernel"], 10, 0)
    dist, ind = NearestNeighbors(n_neighbors=min(knn + 1, n), metric="euclidean").fit(x).kneighbors(x)
    ms = utils.determine_multiscale_space(res)
    ms_t = utils_tight.determine_multiscale_space(
        {"EigenValues": res_t["EigenValues"], "EigenVectors": res_t["EigenVectors"]})
    # pin the oracle
    o = po.run_diffusion_maps(x, 10, knn)
    assert abs(o["kernel"] - res["kernel"]).max() == 0
    assert abs(o["T"] - res["T"]).max() == 0
    assert np.array_equal(o["EigenValues"], res["EigenValues"].values)
    assert np.array_equal(o["EigenVectors"], res["EigenVectors"].values)
    assert np.array_equal(po.multiscale_space(o["EigenVectors"], o["EigenValues"]), ms.values)
    k = _csr_parts(res["kernel"])
    t = _csr_parts(res["T"])
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        x=x, early=info["early"], terminals=np.array(list(info["terminals"].values())),
        knn=knn, knn_dist=dist, knn_ind=ind.astype(np.int32),
        k_indptr=k["indptr"], k_indices=k["indices"], k_data=k["data"], t_data=t["data"],
        evals_asis=res["EigenValues"].values, evecs_asis=res["EigenVectors"].values,
        evals_tight=res_t["EigenValues"].values, evecs_tight=res_t["EigenVectors"].values,
        ms_asis=ms.values, ms_tight=ms_t.values)
    return x, info, names, ms


def case_palantir(core, name, ms, info, names, num_waypoints, knn=30):
    term = {names[v]: k for k, v in info["terminals"].items()}
    pr = core.run_palantir(ms, names[info["early"]], terminal_states=term,
                           num_waypoints=num_waypoints, knn=knn)
    tpos = list(info["terminals"].values())
    o = po.run_palantir(ms.values, info["early"], terminal=tpos, num_waypoints=num_waypoints, knn=knn)
    assert list(names[o["waypoints"]]) == list(pr.waypoints)
    assert np.array_equal(o["pseudotime"], pr.pseudotime.values)
    assert np.array_equal(o["fate_probs"], pr.branch_probs.values)
    assert np.array_equal(o["entropy"], pr.entropy.values)
    assert list(names[o["terminal"]]) == list(pr.branch_probs.columns)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        ms=ms.values, early=info["early"], terminals=np.array(tpos), num_waypoints=num_waypoints, knn=knn,
        waypoints=names.get_indexer(pr.waypoints), pseudotime=pr.pseudotime.values,
        entropy=pr.entropy.values, fate_probs=pr.branch_probs.values,
        fate_probs_raw=o["fate_probs_raw"],     # before presults.py:34's clip: oracle value (pinned above)
        fate_columns=names.get_indexer(pr.branch_probs.columns))


def case_reference_known_answers(core):
    """The two runnable known-answer tests of the reference
    (tests/test_core_equivalence.py:109-132) on their own seeded inputs."""
    rng = np.random.default_rng(0)
    a = pd.DataFrame(rng.normal(size=(30, 3)))
    wp = core._max_min_sampling(a, num_waypoints=9, seed=0)
    assert list(po.max_min_sampling(a.values, 9, 0)) == list(wp)
    rng = np.random.default_rng(1)
    b = pd.DataFrame(rng.normal(size=(25, 4)))
    b.index = pd.Index([f"cell_{i}" for i in range(25)])
    wps = pd.Index([b.index[0], b.index[5], b.index[10], b.index[15]])
    pt, w = core._compute_pseudotime(b, b.index[0], knn=5, waypoints=wps, n_jobs=1, max_iterations=10)
    opt, ow, _ = po.compute_pseudotime(b.values, 0, 5, np.array([0, 5, 10, 15]), 1, 10)
    assert np.array_equal(opt, pt.values) and np.array_equal(ow, w.values)
    np.savez_compressed(os.path.join(OUT, "ref_known_answers.npz"),
                        mm_data=a.values, mm_waypoints=np.asarray(wp, dtype=np.int64),
                        pt_data=b.values, pt_waypoints=np.array([0, 5, 10, 15]),
                        pt_pseudotime=pt.values, pt_w=w.values)


def case_tiny(core, utils):
    """Reference-test-sized input (50 x 10, knn 30 => brute kNN because k+1 >= N//2)."""
    rng = np.random.default_rng(7)
    x = rng.random((50, 10))
    names = pd.Index([f"cell_{i}" for i in range(50)])
    df = pd.DataFrame(x, index=names)
    res = utils.run_diffusion_maps(df, n_components=10, knn=30, kernel_backend="sklearn")
    k = _csr_parts(res["kernel"])
    o = po.run_diffusion_maps(x)
    assert abs(o["kernel"] - res["kernel"]).max() == 0
    np.savez_compressed(os.path.join(OUT, "tiny_50x10.npz"), x=x,
                        k_indptr=k["indptr"], k_indices=k["indices"], k_data=k["data"],
                        evals_asis=res["EigenValues"].values)


def case_wrappers(core, utils, ms, info, names):
    """early_cell / find_terminal_states / fallback_terminal_cell of the reference (utils.py:944-1137) on
    the multiscale matrix of the palantir case, with the generator's components as cell types."""
    ad = utils.AnnData()
    labels = np.array(["trunk", "B0", "B1", "B2"])[info["component"]]
    ad.obs = pd.DataFrame({"celltype": labels}, index=names)
    ad.obs_names = names
    ad.obsm = {"DM_EigenVectors_multiscaled": ms.values}
    ad.obsp, ad.uns = {}, {}
    found = {}
    for ct in ("trunk", "B0", "B1", "B2"):
        try:
            found[ct] = utils.early_cell(ad, ct)
        except utils.CellNotFoundException:
            found[ct] = ""
    ts = utils.find_terminal_states(ad, ["B0", "B1", "B2", "trunk"])
    # a cell type that is at no extreme: the middle third of the trunk
    mid = (info["component"] == 0) & (info["time"] > 0.4) & (info["time"] < 0.6)
    labels2 = labels.copy(); labels2[mid] = "mid"
    ad.obs = pd.DataFrame({"celltype": labels2}, index=names)
    try:
        utils.early_cell(ad, "mid")
        mid_found = True
    except utils.CellNotFoundException:
        mid_found = False
    fb = utils.early_cell(ad, "mid", fallback_seed=7)
    np.savez_compressed(os.path.join(OUT, "wrappers_n1500.npz"), ms=ms.values, labels=labels, labels_mid=labels2,
                        early=np.array([found[c] for c in ("trunk", "B0", "B1", "B2")]),
                        ts_cells=np.asarray(ts.index, dtype=str), ts_types=np.asarray(ts.values, dtype=str),
                        mid_found=mid_found, mid_fallback=str(fb))
    print("wrappers:", found, dict(ts), "mid found directly:", mid_found, "fallback:", fb)


def case_terminal_states(core):
    """identify_terminal_states and the terminal_states=None branch of run_palantir (core.py:404-491, 566-649)
    run by the reference itself.  The reference leaves the ARPACK start of the stationary vector to scipy's
    default, so every case is run twice and must give the same cells both times (the stationary vector is
    unique; only its last bits depend on the start)."""
    out = {}
    cases = []
    rng = np.random.default_rng(0)
    cases.append(("rand5", rng.random((2000, 5)) ** 2, 0, 400))
    rng = np.random.default_rng(2)
    cases.append(("rand6", rng.random((1500, 6)) ** 2, 0, 300))
    x, info = synth_branching(3000, 20, seed=3)
    o = po.run_diffusion_maps(x, 10, 30)
    cases.append(("branch", po.multiscale_space(o["EigenVectors"], o["EigenValues"]), info["early"], 500))
    for tag, ms, early, nwp in cases:
        names = synth_names(ms.shape[0])
        df = pd.DataFrame(ms, index=names)
        runs = [core.identify_terminal_states(df, names[early], knn=30, num_waypoints=nwp) for _ in range(2)]
        assert list(runs[0][0]) == list(runs[1][0]) and list(runs[0][1]) == list(runs[1][1]), tag
        ts, exc = runs[0]
        assert len(ts) > 0, tag
        ot, oe = po.identify_terminal_states(ms, early, knn=30, num_waypoints=nwp)
        assert sorted(names[ot]) == sorted(ts) and sorted(names[oe]) == sorted(exc), tag
        pr = core.run_palantir(df, names[early], terminal_states=None, num_waypoints=nwp, knn=30)
        op = po.run_palantir(ms, early, terminal=None, num_waypoints=nwp, knn=30)
        assert list(names[op["terminal"]]) == list(pr.branch_probs.columns), tag
        assert np.abs(op["pseudotime"] - pr.pseudotime.values).max() == 0, tag
        assert np.abs(op["fate_probs"] - pr.branch_probs.values).max() < 1e-9, tag    # ARPACK start differs
        out.update({f"{tag}_ms": ms, f"{tag}_early": early, f"{tag}_nwp": nwp,
                    f"{tag}_terminal": names.get_indexer(pd.Index(ts)), f"{tag}_excluded": names.get_indexer(exc),
                    f"{tag}_pseudotime": pr.pseudotime.values, f"{tag}_fate_probs": pr.branch_probs.values,
                    f"{tag}_fate_columns": names.get_indexer(pr.branch_probs.columns),
                    f"{tag}_entropy": pr.entropy.values})
        print("terminal states", tag, list(ts), "excluded", len(exc))
    np.savez_compressed(os.path.join(OUT, "terminal_states.npz"), **out)


def main():
    warnings.simplefilter("ignore")
    os.makedirs(OUT, exist_ok=True)
    core, utils = ref_shim.load_reference()
    _, utils_tight = ref_shim.load_reference(eig_tol=1e-12)
    x, info, names, ms = case_diffusion(core, utils, utils_tight, "dm_f32_n1500_d20", 1500, 20, np.float32)
    case_palantir(core, "palantir_n1500_wp200", ms, info, names, 200)
    case_wrappers(core, utils, ms, info, names)
    case_diffusion(core, utils, utils_tight, "dm_f64_n600_d8", 600, 8, np.float64)   # d <= 15: kd-tree distances
    case_diffusion(core, utils, utils_tight, "dm_f64_n800_d24", 800, 24, np.float64)  # f64 brute
    case_reference_known_answers(core)
    case_tiny(core, utils)
    case_terminal_states(core)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
