"""GPU parity tests: the CUDA path (reached through the C-ABI library) against the CPU
oracle on the same seeded inputs, against the reference outputs committed under
tests/golden/, and -- at benchmark sizes -- through size-independent properties.

Tolerances (BASELINE.json north_star): kNN indices exact (sets per row; ties excepted),
kernel values <= 1e-12 relative, eigenvalues <= 1e-5 relative, eigenvector subspace angle
< 1e-4 (vs the converged oracle), pseudotime max-abs <= 1e-4 and Spearman > 0.999, fate
probabilities max-abs <= 1e-4.
"""
import warnings

import numpy as np
import pandas as pd
import pytest
from scipy.linalg import subspace_angles
from scipy.sparse import csr_matrix
from scipy.stats import spearmanr

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import palantir_b200
    from palantir_b200 import _native

    _native.load()
    return palantir_b200


@pytest.fixture(scope="module")
def po():
    from oracle import palantir_oracle

    return palantir_oracle


def _knn(pb, x, k):
    from palantir_b200 import _device as dv

    idx, dist, stats = dv.knn_like_sklearn(dv.to_device(np.ascontiguousarray(x)), k)
    return idx.cpu().numpy(), dist.cpu().numpy(), stats


def _rows_equal_as_sets(a, b):
    return sum(set(r) == set(s) for r, s in zip(a, b))


# ------------------------------------------------------------------ kNN
@pytest.mark.parametrize("name,rtol", [("dm_f32_n1500_d20.npz", 0.0), ("dm_f64_n800_d24.npz", 1e-12),
                                       ("dm_f64_n600_d8.npz", 0.0)])
def test_knn_matches_reference_fixture(pb, name, rtol):
    z = golden(name)
    idx, dist, stats = _knn(pb, z["x"], int(z["knn"]) + 1)
    assert _rows_equal_as_sets(idx, z["knn_ind"]) == idx.shape[0]
    got, ref = np.sort(dist, 1)[:, 1:], np.sort(z["knn_dist"], 1)[:, 1:]      # column 0 = self (rounding noise)
    if rtol == 0.0:
        assert np.array_equal(got, ref)          # f32 input: sqrtf-quantised, bit-equal; kd-tree case: exact
    else:
        assert np.max(np.abs(got - ref) / ref) <= rtol


@pytest.mark.parametrize("n,d,k,dtype", [
    (50, 10, 31, np.float64),      # reference-test size: k+1 >= N//2 -> brute, N < one tile
    (257, 16, 31, np.float32),     # ragged: one row past a 256 tile, d = one chunk
    (1000, 50, 31, np.float32),    # d with a partial last chunk (48 + 4 -> tail)
    (3000, 100, 31, np.float32),   # config-3 dimensionality
    (777, 130, 11, np.float64),    # d > 128, small k
    (2000, 3, 30, np.float64),     # kd-tree rule (d <= 15): direct differences
    (600, 20, 50, np.float32),     # k wider than the candidate buffers -> exact path
])
def test_knn_matches_sklearn_live(pb, po, n, d, k, dtype):
    rng = np.random.default_rng(n + d)
    x = (rng.normal(size=(n, d)) * rng.uniform(0.2, 3.0, size=d)).astype(dtype)
    ref_d, ref_i = po.knn_with_self(x, k)
    idx, dist, stats = _knn(pb, x, k)
    bad = [r for r in range(n) if set(idx[r]) != set(ref_i[r])]
    # a mismatch is only admissible at an exact / rounding-level tie of the k-th distance
    for r in bad:
        assert abs(np.sort(dist[r])[-1] - np.sort(ref_d[r])[-1]) <= 1e-12 * max(1.0, ref_d[r].max())
    assert len(bad) <= max(1, n // 500)
    tol = 0.0 if dtype == np.float32 and d > 15 else 1e-11
    a, b = np.sort(dist, 1)[:, 1:], np.sort(ref_d, 1)[:, 1:]
    assert np.max(np.abs(a - b)) <= tol * max(1.0, b.max()) + (0 if tol else 0)


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("n,d,dtype", [(3000, 100, np.float32), (1500, 20, np.float32), (2500, 128, np.float64),
                                       (700, 17, np.float32), (5000, 50, np.float32)])
def test_knn_both_engines_match_sklearn(pb, po, engine, n, d, dtype):
    """FP32 SIMT kernel and tcgen05 3xTF32 kernel: identical neighbour sets, sklearn's distances."""
    from palantir_b200 import _device as dv

    rng = np.random.default_rng(n * 7 + d)
    x = (rng.normal(size=(n, d)) * rng.uniform(0.3, 2.0, size=d) + rng.normal(size=d)).astype(dtype)
    ref_d, ref_i = po.knn_with_self(x, 31)
    old = dv.KNN_ENGINE
    dv.KNN_ENGINE = engine
    try:
        idx, dist, stats = _knn(pb, x, 31)
    finally:
        dv.KNN_ENGINE = old
    assert stats["engine"] == engine
    assert _rows_equal_as_sets(idx, ref_i) == n
    a, b = np.sort(dist, 1)[:, 1:], np.sort(ref_d, 1)[:, 1:]
    if dtype == np.float32:
        assert np.array_equal(a, b)
    else:
        assert np.max(np.abs(a - b) / b) <= 1e-12


def test_knn_with_duplicate_points(pb, po):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(400, 24)).astype(np.float32)
    x[100:110] = x[0:10]                      # exact duplicates: zero distances, ties
    ref_d, ref_i = po.knn_with_self(x, 16)
    idx, dist, _ = _knn(pb, x, 16)
    assert np.array_equal(np.sort(dist, 1)[:, 2:], np.sort(ref_d, 1)[:, 2:])
    # every returned neighbour is a true k-NN (distance not larger than the reference's k-th)
    assert np.all(np.sort(dist, 1)[:, -1] <= np.sort(ref_d, 1)[:, -1])


# ------------------------------------------------------------------ kernel, T, eigensolver
@pytest.mark.parametrize("name", ["dm_f32_n1500_d20.npz", "dm_f64_n600_d8.npz", "dm_f64_n800_d24.npz"])
def test_diffusion_maps_match_reference_fixture(pb, name):
    z = golden(name)
    res = pb.utils.run_diffusion_maps(pd.DataFrame(z["x"]), n_components=10, knn=int(z["knn"]),
                                      kernel_backend="sklearn")
    k = res["kernel"]
    assert isinstance(k, csr_matrix) and set(res) == {"T", "EigenVectors", "EigenValues", "kernel"}
    assert np.array_equal(k.indptr, z["k_indptr"]) and np.array_equal(k.indices, z["k_indices"])
    assert np.max(np.abs(k.data - z["k_data"]) / z["k_data"]) <= 1e-12
    assert np.max(np.abs(res["T"].data - z["t_data"]) / z["t_data"]) <= 1e-12
    ev = res["EigenValues"].values
    assert np.max(np.abs(ev - z["evals_tight"]) / np.abs(z["evals_tight"])) <= 1e-5
    assert np.max(np.abs(ev - z["evals_asis"])) <= 1e-4          # the reference's own test tolerance
    v = res["EigenVectors"].values
    assert subspace_angles(v, z["evecs_tight"]).max() < 1e-4
    assert np.allclose(np.linalg.norm(v, axis=0), 1.0, atol=1e-12)
    assert np.min(np.abs(np.sum(v * z["evecs_tight"], axis=0))) > 1 - 1e-8      # equal up to sign, per vector


def test_kernel_alpha_and_tiny_input(pb, po):
    z = golden("tiny_50x10.npz")
    k = pb.utils.compute_kernel(pd.DataFrame(z["x"]), backend="sklearn")
    assert np.array_equal(k.indices, z["k_indices"])
    assert np.max(np.abs(k.data - z["k_data"]) / z["k_data"]) <= 1e-12
    rng = np.random.default_rng(2)
    x = rng.normal(size=(700, 20))
    ka = pb.utils.compute_kernel(pd.DataFrame(x), knn=10, alpha=0.5, backend="sklearn")
    ko = po.adaptive_kernel(x, 10, 0.5)
    ko.sort_indices()
    assert np.array_equal(ka.indices, ko.indices)
    assert np.max(np.abs(ka.data - ko.data) / ko.data) <= 1e-12


def test_diffusion_maps_from_kernel_reference_tests(pb):
    """tests/test_utils_diffusion_maps_from_kernel.py of the reference, same assertions."""
    from scipy.sparse.linalg import eigs

    rng = np.random.default_rng(0)
    a = rng.random((50, 50))
    kern = csr_matrix((a + a.T) / 2)
    r = pb.utils.diffusion_maps_from_kernel(kern)
    assert r["T"].shape == (50, 50) and r["EigenVectors"].shape == (50, 10) and r["EigenValues"].shape == (10,)
    assert pb.utils.diffusion_maps_from_kernel(kern, n_components=5)["EigenVectors"].shape == (50, 5)
    r2 = pb.utils.diffusion_maps_from_kernel(kern, seed=0)
    assert np.allclose(r["EigenValues"], r2["EigenValues"])
    e, _ = eigs(r["T"].toarray(), 10, tol=1e-4, maxiter=1000)
    assert np.allclose(r["EigenValues"], np.real(sorted(e, reverse=True)[:10]), atol=1e-4)
    assert pb.utils.determine_multiscale_space(r).shape[0] == 50


def test_kernel_that_is_not_symmetric(pb):
    """diffusion_maps_from_kernel accepts any kernel (utils.py:477-484 hands T to ARPACK's general solver): a
    directed kNN affinity matrix goes through the Arnoldi path; real parts of the leading eigenpairs as in the
    reference (utils.py:485-495), compared with scipy's eigs run to convergence."""
    from scipy.sparse.linalg import eigs
    from sklearn.neighbors import NearestNeighbors

    rng = np.random.default_rng(8)
    pts = rng.normal(size=(700, 3)) * np.array([4.0, 1.0, 0.3])
    g = NearestNeighbors(n_neighbors=12).fit(pts).kneighbors_graph(pts, mode="distance")
    g.data = np.exp(-g.data / np.median(g.data))              # W only, no W + W^T: not symmetric
    k = csr_matrix(g)
    assert abs(k - k.T).max() > 0.1
    res = pb.utils.diffusion_maps_from_kernel(k, n_components=6, seed=0)
    assert pb.utils.LAST_INFO["lanczos"]["solver"] == "arnoldi"
    deg = np.asarray(k.sum(axis=1)).ravel()
    t = csr_matrix((1.0 / deg)[:, None] * k.toarray())
    assert abs(res["T"] - t).max() < 1e-14
    w, v = eigs(t, 6, tol=1e-13, maxiter=20000, v0=np.ones(700))
    order = np.argsort(-np.real(w))
    w, v = np.real(w[order]), np.real(v[:, order])
    assert np.max(np.abs(res["EigenValues"].values - w)) < 1e-8
    got = res["EigenVectors"].values
    assert np.allclose(np.linalg.norm(got, axis=0), 1.0)
    for c in range(6):
        if np.abs(np.imag(eigs(t, 6, tol=1e-13, maxiter=20000, v0=np.ones(700))[0][order][c])) > 0:
            continue                                           # complex pair: the real part of a vector is not unique
        ref = v[:, c] / np.linalg.norm(v[:, c])
        assert min(np.abs(got[:, c] - ref).max(), np.abs(got[:, c] + ref).max()) < 1e-6


def test_kernel_edited_in_place_is_uploaded_again(pb):
    """compute_kernel leaves a device copy on the matrix it returns; editing the matrix in place must
    not let diffusion_maps_from_kernel use the stale copy."""
    z = golden("dm_f64_n600_d8.npz")
    kernel = pb.utils.compute_kernel(pd.DataFrame(z["x"]), knn=int(z["knn"]))
    kernel.data[::2] *= 0.25
    sym = csr_matrix(kernel.maximum(kernel.T))          # fresh symmetric matrix, no device copy attached
    kernel.data[:] = sym.data                           # same pattern: edit in place to the symmetric values
    got = pb.utils.diffusion_maps_from_kernel(kernel, n_components=5, seed=0)
    want = pb.utils.diffusion_maps_from_kernel(sym, n_components=5, seed=0)
    assert np.allclose(got["EigenValues"].values, want["EigenValues"].values, rtol=0, atol=1e-12)
    assert abs(got["T"] - want["T"]).max() < 1e-15


def test_device_copies_are_off_by_default(pb):
    """config.REUSE_DEVICE_RESULTS defaults to False: every call uploads the host arrays it is given, so an
    in-place edit of a single entry between two calls is always honoured."""
    from palantir_b200 import _devcache

    assert pb.config.REUSE_DEVICE_RESULTS is False
    z = golden("dm_f64_n800_d24.npz")
    res = pb.utils.run_diffusion_maps(pd.DataFrame(z["x"]), knn=int(z["knn"]), kernel_backend="sklearn")
    assert _devcache.lookup(res["EigenVectors"].values) is None
    ms = pb.utils.determine_multiscale_space(res)
    assert _devcache.lookup(ms.values) is None
    early = ms.index[int(z["early"])]
    a = pb.core.run_palantir(ms, early, num_waypoints=100, knn=20)
    ms2 = ms.copy()
    ms2.iloc[5, 1] += 0.25 * float(ms.iloc[:, 1].std())          # one entry
    b = pb.core.run_palantir(ms2, early, num_waypoints=100, knn=20)
    assert not np.array_equal(a.pseudotime.values, b.pseudotime.values)


def test_device_copies_are_equivalent_and_never_stale(pb, monkeypatch):
    """Opt-in (config.REUSE_DEVICE_RESULTS = True): results handed from call to call keep a device copy
    (palantir_b200/_devcache.py); using it must give the same numbers as the host arrays, and an array edited
    in place as a whole must not be served from the copy."""
    from palantir_b200 import _devcache

    monkeypatch.setattr(pb.config, "REUSE_DEVICE_RESULTS", True)
    z = golden("dm_f64_n800_d24.npz")
    res = pb.utils.run_diffusion_maps(pd.DataFrame(z["x"]), knn=int(z["knn"]), kernel_backend="sklearn")
    assert _devcache.lookup(res["EigenVectors"].values) is not None
    ms_fast = pb.utils.determine_multiscale_space(res)
    plain = {"EigenValues": res["EigenValues"].copy(), "EigenVectors": pd.DataFrame(res["EigenVectors"].values.copy())}
    assert _devcache.lookup(plain["EigenVectors"].values) is None
    ms_host = pb.utils.determine_multiscale_space(plain)
    assert np.array_equal(ms_fast.values, ms_host.values)
    # downstream: run_palantir from the cached multiscale matrix == from a plain copy of it
    early = ms_fast.index[int(z["early"])]
    a = pb.core.run_palantir(ms_fast, early, num_waypoints=100, knn=20)
    b = pb.core.run_palantir(pd.DataFrame(ms_fast.values.copy(), index=ms_fast.index), early, num_waypoints=100, knn=20)
    assert np.array_equal(a.pseudotime.values, b.pseudotime.values)
    assert np.array_equal(a.branch_probs.values, b.branch_probs.values)
    # in-place edit of the eigenvectors: the stale device copy must not be used
    vals = res["EigenVectors"].values
    if not vals.flags.writeable:
        vals = res["EigenVectors"]._mgr.blocks[0].values.T if hasattr(res["EigenVectors"], "_mgr") else vals
    if vals.flags.writeable:
        vals[:, 1] *= -1.0
        ms_edit = pb.utils.determine_multiscale_space(res)
        assert np.array_equal(ms_edit.values[:, 0], -ms_fast.values[:, 0])
        assert np.array_equal(ms_edit.values[:, 1:], ms_fast.values[:, 1:])


def _wrapper_adata(pb, labels_key="labels"):
    from palantir_b200.compat import AnnDataLike
    from palantir_b200.datasets import synth_names

    z = golden("wrappers_n1500.npz")
    ad = AnnDataLike(X=np.zeros((1500, 1)), obs_names=synth_names(1500))
    ad.obs["celltype"] = z[labels_key]
    ad.obsm["DM_EigenVectors_multiscaled"] = z["ms"]
    return ad, z


def test_early_cell_and_find_terminal_states_match_reference(pb):
    """utils.py:944-1137 of the reference on the fixture it produced (oracle/make_golden.py::case_wrappers)."""
    ad, z = _wrapper_adata(pb)
    for ct, want in zip(("trunk", "B0", "B1", "B2"), z["early"]):
        assert pb.utils.early_cell(ad, ct) == str(want)
    ts = pb.utils.find_terminal_states(ad, ["B0", "B1", "B2", "trunk"])
    assert list(ts.index) == [str(c) for c in z["ts_cells"]] and list(ts.values) == [str(t) for t in z["ts_types"]]


def test_early_cell_fallback_matches_reference(pb):
    ad, z = _wrapper_adata(pb, "labels_mid")
    with pytest.raises(pb.utils.CellNotFoundException):
        pb.utils.early_cell(ad, "mid")
    with pytest.warns(UserWarning, match="will be skipped"):
        ts = pb.utils.find_terminal_states(ad, ["mid", "B0"])
    assert list(ts.values) == ["B0"]
    assert pb.utils.early_cell(ad, "mid", fallback_seed=7) == str(z["mid_fallback"])
    assert "palantir_pseudotime" in ad.obs.columns          # the fallback run writes into `ad`, as the reference's does


def test_lanczos_basis_growth_gives_the_same_eigenpairs(pb):
    """The Lanczos basis is allocated for 320 vectors and re-allocated when more are needed."""
    from palantir_b200 import _device as dv

    z = golden("dm_f64_n800_d24.npz")
    df = pd.DataFrame(z["x"])
    a = pb.utils.run_diffusion_maps(df, knn=int(z["knn"]), kernel_backend="sklearn")
    old = dv.LANCZOS_INITIAL_BASIS
    dv.LANCZOS_INITIAL_BASIS = 1
    try:
        b = pb.utils.run_diffusion_maps(df, knn=int(z["knn"]), kernel_backend="sklearn")
    finally:
        dv.LANCZOS_INITIAL_BASIS = old
    assert pb.utils.LAST_INFO["lanczos"]["steps"] > 20            # grew at least once (first allocation: one batch)
    assert np.array_equal(a["EigenValues"].values, b["EigenValues"].values)
    assert np.array_equal(a["EigenVectors"].values, b["EigenVectors"].values)


def test_run_diffusion_maps_anndata_keys(pb):
    from palantir_b200.compat import AnnDataLike

    z = golden("dm_f64_n600_d8.npz")
    ad = AnnDataLike(np.zeros((600, 2)))
    ad.obsm["X_pca"] = z["x"]
    res = pb.utils.run_diffusion_maps(ad, n_components=10, kernel_backend="sklearn")
    assert ad.obsp["DM_Kernel"] is res["kernel"] and ad.obsp["DM_Similarity"] is res["T"]
    assert np.array_equal(ad.obsm["DM_EigenVectors"], res["EigenVectors"].values)
    assert np.array_equal(ad.uns["DM_EigenValues"], res["EigenValues"].values)
    k2 = pb.utils.compute_kernel(ad, knn=10, alpha=0.5, kernel_key="custom")
    assert "custom" in ad.obsp and isinstance(k2, csr_matrix)
    with pytest.warns(UserWarning):
        pb.utils.run_diffusion_maps(res["kernel"])            # sparse input = kernel


# ------------------------------------------------------------------ run_palantir
@pytest.mark.parametrize("n,m,k", [(5000, 5, 30), (1000, 3, 10), (777, 1, 5), (4100, 9, 31), (300, 2, 30)])
def test_boxed_low_dimensional_knn_matches_sklearn_kd_tree(pb, n, m, k):
    """The multiscale-space search (core.py:355-356): Morton order + bounding-box pruning must return
    sklearn's kd-tree neighbours with bit-equal distances (direct differences in FP64)."""
    import torch
    from sklearn.neighbors import NearestNeighbors
    from palantir_b200 import _device as dv, _device_core as dc

    rng = np.random.default_rng(n + m)
    t = rng.random(n)
    data = np.stack([np.sin((c + 1) * t) + 0.01 * rng.normal(size=n) for c in range(m)], 1)   # a noisy curve
    dev = dv.to_device(np.ascontiguousarray(data))
    perm, inv, work = dc.morton_order(dev)
    idx, dist, st = dc.knn_table_boxed(work, k)
    perm_h = perm.cpu().numpy()
    assert sorted(perm_h.tolist()) == list(range(n))
    got_idx = np.empty((n, k), np.int64)
    got_dist = np.empty((n, k))
    got_idx[perm_h] = perm_h[idx.cpu().numpy()]
    got_dist[perm_h] = dist.cpu().numpy()
    rd, ri = NearestNeighbors(n_neighbors=k, algorithm="kd_tree").fit(data).kneighbors(data)
    assert np.array_equal(got_dist, rd)
    rows_equal = (np.sort(got_idx, 1) == np.sort(ri, 1)).all(1)
    ties = np.array([len(np.unique(r)) < k for r in rd])
    assert np.all(rows_equal | ties)
    assert int(st["scanned"].item()) <= ((n + 127) // 128) ** 2


def test_max_min_sampling_known_answer(pb):
    z = golden("ref_known_answers.npz")
    wp = pb.core._max_min_sampling(pd.DataFrame(z["mm_data"]), num_waypoints=9, seed=0)
    assert list(wp) == list(z["mm_waypoints"])


def test_compute_pseudotime_known_answer(pb):
    """tests/test_core_equivalence.py:117-132 of the reference (atol = rtol = 1e-10)."""
    z = golden("ref_known_answers.npz")
    b = pd.DataFrame(z["pt_data"], index=[f"cell_{i}" for i in range(25)])
    pt, w = pb.core._compute_pseudotime(b, "cell_0", knn=5, waypoints=b.index[z["pt_waypoints"]], n_jobs=1,
                                        max_iterations=10)
    assert np.allclose(pt.values, z["pt_pseudotime"], atol=1e-10, rtol=1e-10)
    assert np.allclose(w.values, z["pt_w"], atol=1e-10, rtol=1e-10)


def test_run_palantir_matches_reference_fixture(pb, names1500):
    z = golden("palantir_n1500_wp200.npz")
    names = names1500
    ms = pd.DataFrame(z["ms"], index=names)
    pr = pb.core.run_palantir(ms, names[int(z["early"])], terminal_states=list(names[z["terminals"]]),
                              num_waypoints=int(z["num_waypoints"]), knn=int(z["knn"]))
    assert isinstance(pr, pb.PResults)
    assert list(names.get_indexer(pr.waypoints)) == list(z["waypoints"])          # identical waypoints
    assert np.max(np.abs(pr.pseudotime.values - z["pseudotime"])) <= 1e-4
    assert spearmanr(pr.pseudotime.values, z["pseudotime"])[0] > 0.999
    assert list(names.get_indexer(pr.branch_probs.columns)) == list(z["fate_columns"])
    assert np.max(np.abs(pr.branch_probs.values - z["fate_probs"])) <= 1e-4
    assert np.max(np.abs(pr.entropy.values - z["entropy"])) <= 1e-4
    # much tighter in practice (FP64 shortest paths / refinement / solve)
    assert np.max(np.abs(pr.pseudotime.values - z["pseudotime"])) <= 1e-10


def test_run_palantir_vs_oracle_other_settings(pb, po, names1500):
    z = golden("palantir_n1500_wp200.npz")
    names, ms = names1500, z["ms"]
    for kw in (dict(scale_components=False), dict(use_early_cell_as_start=True), dict(knn=12, num_waypoints=60)):
        o = po.run_palantir(ms, int(z["early"]), terminal=list(z["terminals"]), **{"num_waypoints": 120, **kw})
        pr = pb.core.run_palantir(pd.DataFrame(ms, index=names), names[int(z["early"])],
                                  terminal_states=list(names[z["terminals"]]), **{"num_waypoints": 120, **kw})
        assert list(names.get_indexer(pr.waypoints)) == list(o["waypoints"]), kw
        assert np.max(np.abs(pr.pseudotime.values - o["pseudotime"])) <= 1e-9, kw
        assert np.max(np.abs(pr.branch_probs.values - o["fate_probs"])) <= 1e-9, kw


def test_run_palantir_reference_api_behaviour(pb):
    """tests/test_core_run_palantir.py of the reference (sizes and assertions)."""
    from palantir_b200.compat import AnnDataLike

    rng = np.random.default_rng(11)
    df = pd.DataFrame(rng.random((50, 10)), columns=[f"c{i}" for i in range(10)],
                      index=[f"cell_{i}" for i in range(50)])
    pr = pb.core.run_palantir(df, "cell_0")                                   # terminal states auto-detected
    assert isinstance(pr, pb.PResults) and len(pr.pseudotime) == 50
    ad = AnnDataLike(np.zeros((50, 3)), obs_names=df.index)
    ad.obsm["DM_EigenVectors_multiscaled"] = df.values
    pb.core.run_palantir(ad, "cell_0")
    for key in ("palantir_pseudotime", "palantir_entropy"):
        assert key in ad.obs
    assert "palantir_fate_probabilities" in ad.obsm and "palantir_waypoints" in ad.uns
    pr2 = pb.core.run_palantir(df, "cell_0", terminal_states=["cell_1", "cell_2"])
    assert "cell_1" in pr2.branch_probs.columns and "cell_2" in pr2.branch_probs.columns
    a = pb.core.run_palantir(df, "cell_0", terminal_states=["cell_1", "cell_2"], scale_components=True)
    b = pb.core.run_palantir(df, "cell_0", terminal_states=["cell_1", "cell_2"], scale_components=False)
    assert not np.array_equal(a.pseudotime.values, b.pseudotime.values)
    with pytest.raises(ValueError):
        pb.core.run_palantir(df, "cell_0", knn=0)
    ts = pd.Series({"cell_1": "A", "cell_2": "B"})
    pb.core.run_palantir(ad, "cell_0", terminal_states=ts, save_as_df=False)
    assert list(ad.uns["palantir_fate_probabilities_columns"]) == ["A", "B"]


@pytest.mark.parametrize("tag", ["rand5", "rand6", "branch"])
def test_terminal_states_match_reference_fixture(pb, tag):
    """core.identify_terminal_states / _terminal_states_from_markov_chain and the terminal_states=None branch
    of run_palantir (reference core.py:404-491, 566-649) against outputs of the reference's own functions
    (tests/golden/terminal_states.npz).  The device power iteration replaces ARPACK's randomly started k=1 solve:
    the stationary vector is unique, so cells must be identical and fates agree to solver accuracy."""
    from palantir_b200.datasets import synth_names

    z = golden("terminal_states.npz")
    ms, early, nwp = z[f"{tag}_ms"], int(z[f"{tag}_early"]), int(z[f"{tag}_nwp"])
    names = synth_names(ms.shape[0])
    df = pd.DataFrame(ms, index=names)
    term, exc = pb.core.identify_terminal_states(df, names[early], knn=30, num_waypoints=nwp)
    assert sorted(names.get_indexer(pd.Index(term))) == sorted(z[f"{tag}_terminal"])
    assert sorted(names.get_indexer(exc)) == sorted(z[f"{tag}_excluded"])
    pr = pb.core.run_palantir(df, names[early], terminal_states=None, num_waypoints=nwp, knn=30)
    assert list(names.get_indexer(pr.branch_probs.columns)) == list(z[f"{tag}_fate_columns"])
    assert np.max(np.abs(pr.pseudotime.values - z[f"{tag}_pseudotime"])) <= 1e-10
    assert np.max(np.abs(pr.branch_probs.values - z[f"{tag}_fate_probs"])) <= 1e-8
    assert np.max(np.abs(pr.entropy.values - z[f"{tag}_entropy"])) <= 1e-8


def test_stationary_ranks_vs_oracle_chain(pb, po):
    """_terminal_states_from_markov_chain (core.py:590-649) on the oracle's own waypoint chain: same positions."""
    from palantir_b200 import _device as dv

    z = golden("terminal_states.npz")
    ms, early, nwp = z["rand5_ms"], int(z["rand5_early"]), int(z["rand5_nwp"])
    data = po.minmax_scale(ms)
    bnd = po.boundary_cells(data)
    start = po.pick_start_cell(data, bnd, early)
    wp = po.assemble_waypoints(po.max_min_sampling(data, nwp, 20), bnd, None, start)
    pt, _, _ = po.compute_pseudotime(data, start, 30, wp)
    chain = po.markov_chain(data[wp], 30, pt[wp])
    want = po.terminal_states_from_chain(chain, data[wp], pt[wp])
    labels = pd.Index(wp)
    got = pb.core._terminal_states_from_markov_chain(dv.to_device(np.ascontiguousarray(chain.toarray())),
                                                     dv.to_device(np.ascontiguousarray(data[wp])), labels, pt[wp])
    assert sorted(labels.get_indexer(pd.Index(got))) == sorted(want)


def test_missing_labels_raise_keyerror(pb, names1500):
    """A waypoint / terminal label that is not a cell of the data raises KeyError like the reference's
    data.loc[waypoints] (core.py:209, 355) instead of reaching the device as position -1."""
    z = golden("palantir_n1500_wp200.npz")
    names = names1500
    ms = pd.DataFrame(z["ms"], index=names)
    early = names[int(z["early"])]
    with pytest.raises(KeyError):
        pb.core.run_palantir(ms, early, terminal_states=[names[5], "no_such_cell"], num_waypoints=60)
    with pytest.raises(KeyError):
        pb.core.run_palantir(ms, early, num_waypoints=pd.Index([names[3], names[9], "ghost"]))
    with pytest.raises(KeyError):
        pb.core._compute_pseudotime(ms, early, knn=10, waypoints=pd.Index([early, "ghost"]), n_jobs=1)


# ------------------------------------------------------------------ stage kernels
def _random_knn_graph(n, m, k, seed):
    from sklearn.neighbors import NearestNeighbors

    rng = np.random.default_rng(seed)
    data = rng.random((n, m))
    adj = NearestNeighbors(n_neighbors=k).fit(data).kneighbors_graph(data, mode="distance")
    return data, adj


def test_sssp_bit_exact_vs_scipy_dijkstra(pb):
    from scipy.sparse import csgraph

    from palantir_b200 import _device as dv, _device_core as dc

    data, adj = _random_knn_graph(4000, 3, 8, 0)
    idx, dist, _ = dc.knn_table(dv.to_device(data), 8)
    g = dc.undirected_graph(idx, dist)
    src = np.array([0, 17, 3999, 1234, 2048], dtype=np.int64)
    for delta in (None, 1e-3, 1e9):
        D, stats = dc.sssp(g, dv.to_device(src), delta=delta, want_stats=True)
        ref = csgraph.dijkstra(adj, directed=False, indices=src)
        assert np.array_equal(D.cpu().numpy(), ref)                  # +inf where unreachable, bit-equal elsewhere


def test_graph_components_vs_scipy_repeatable(pb):
    """Reachability labels (core.py:806-817) on a graph with deep trees and a few components, five
    times: the labels must be the component's smallest index every time (a racing path-halving store
    once left inner tree nodes as labels for a handful of cells per million)."""
    import torch
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from palantir_b200 import _device as dv

    rng = np.random.default_rng(5)
    n = 400_000
    cuts = np.sort(rng.choice(np.arange(1, n), 6, replace=False))
    a = np.arange(n - 1)
    keep = ~np.isin(a + 1, cuts)                                  # paths 0..c1-1, c1..c2-1, ...
    src, dst = a[keep], a[keep] + 1
    comp_of = np.searchsorted(cuts, np.arange(n), side="right")
    e1 = rng.integers(0, n, 3 * n)
    e2 = rng.integers(0, n, 3 * n)
    same = comp_of[e1] == comp_of[e2]                              # random chords inside a component
    src, dst = np.concatenate([src, e1[same]]), np.concatenate([dst, e2[same]])
    m = coo_matrix((np.ones(2 * len(src)), (np.concatenate([src, dst]), np.concatenate([dst, src]))), shape=(n, n)).tocsr()
    m.sum_duplicates()
    ncomp, ref = connected_components(m, directed=False)
    assert ncomp == 7
    first = np.full(ncomp, n, dtype=np.int64)
    np.minimum.at(first, ref, np.arange(n))
    want = first[ref]
    indptr, indices = dv.to_device(m.indptr.astype(np.int32)), dv.to_device(m.indices.astype(np.int32))
    label = dv.empty((n,), dv.I32)
    for _ in range(5):
        dv.call("pb200_graph_components", indptr, indices, n, label)
        assert np.array_equal(label.cpu().numpy(), want)


def test_connect_graph_reconnects_like_the_reference(pb, po):
    rng = np.random.default_rng(4)
    a = rng.normal(size=(300, 4))
    b = rng.normal(size=(40, 4)) + 25.0                  # a far-away island: unreachable with knn = 5
    x = np.vstack([a, b])
    names = pd.Index([f"c{i:04d}" for i in range(x.shape[0])])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = po.run_palantir(x, 0, terminal=[50, 320], num_waypoints=40, knn=5, scale_components=False)
        with pytest.warns(UserWarning):
            pr = pb.core.run_palantir(pd.DataFrame(x, index=names), names[0], terminal_states=[names[50], names[320]],
                                      num_waypoints=40, knn=5, scale_components=False)
    assert list(names.get_indexer(pr.waypoints)) == list(o["waypoints"])
    assert np.all(np.isfinite(pr.pseudotime.values))
    assert np.max(np.abs(pr.pseudotime.values - o["pseudotime"])) <= 1e-9
    assert np.max(np.abs(pr.branch_probs.values - o["fate_probs"])) <= 1e-9


def test_markov_chain_and_absorption_vs_oracle(pb, po):
    from palantir_b200 import _device as dv, _device_core as dc

    rng = np.random.default_rng(8)
    wp = rng.random((500, 5))
    pt = rng.random(500)
    T = dc.markov_chain_dense(dv.to_device(wp), 30, dv.to_device(pt)).cpu().numpy()
    To = po.markov_chain(wp, 30, pt).toarray()
    assert np.max(np.abs(T - To)) <= 1e-13
    abs_states = np.array([3, 77, 401])
    B = dc.absorption_probabilities(dv.to_device(To), abs_states).cpu().numpy()
    rows, _, Bo = po.absorption_probabilities(csr_matrix(To), abs_states)
    full = np.empty_like(Bo)
    full[rows] = Bo
    assert np.max(np.abs(B - full)) <= 1e-11


def test_dense_solve_vs_numpy(pb):
    from palantir_b200 import _device as dv

    rng = np.random.default_rng(1)
    for n, r in ((1, 1), (7, 3), (300, 5), (1201, 3)):
        a = rng.normal(size=(n, n)) + n * np.eye(n) * 0.05
        b = rng.normal(size=(n, r))
        ad, bd = dv.to_device(a), dv.to_device(b)
        sing = dv.empty((1,), dv.I32)
        dv.call("pb200_dense_solve_f64", ad, bd, n, r, sing)
        x = np.linalg.solve(a, b)
        assert int(sing.item()) == 0
        assert np.max(np.abs(bd.cpu().numpy() - x)) <= 1e-9 * max(1.0, np.abs(x).max())


def test_spmv_vs_scipy(pb):
    from palantir_b200 import _device as dv

    z = golden("dm_f32_n1500_d20.npz")
    k = csr_matrix((z["k_data"], z["k_indices"], z["k_indptr"]), shape=(1500, 1500))
    kd = dv.csr_from_host(k)
    x = np.random.default_rng(0).normal(size=1500)
    y = dv.empty((1500,), dv.F64)
    dv.call("pb200_csr_spmv_f64", kd.indptr, kd.indices, kd.data, 1500, kd.nnz, dv.to_device(x), y)
    assert np.max(np.abs(y.cpu().numpy() - k @ x)) <= 1e-12


def test_spmv_with_hub_rows_vs_scipy(pb):
    """Rows above 128 entries are walked by a whole warp (the symmetrised kNN graph has hubs of thousands)."""
    from scipy.sparse import random as sprandom, vstack
    from palantir_b200 import _device as dv, _native

    rng = np.random.default_rng(3)
    n = 20_000
    a = sprandom(n, n, density=40 / n, random_state=3, format="lil")
    for r, ln in ((5, 129), (77, 1000), (n - 1, 5000), (1234, 128)):
        cols = rng.choice(n, ln, replace=False)
        a.rows[r] = sorted(cols.tolist()); a.data[r] = rng.normal(size=ln).tolist()
    a = a.tocsr()
    x = rng.normal(size=n)
    lens = np.diff(a.indptr)
    long_rows = np.flatnonzero(lens > int(_native.load().pb200_spmv_long_row())).astype(np.int32)
    assert len(long_rows) >= 3 and 1234 not in long_rows
    y = dv.empty((n,), dv.F64)
    dv.call("pb200_csr_spmv_split_f64", dv.to_device(a.indptr.astype(np.int32)), dv.to_device(a.indices.astype(np.int32)),
            dv.to_device(a.data), n, a.nnz, dv.to_device(long_rows), len(long_rows), dv.to_device(x), y)
    ref = a @ x
    assert np.max(np.abs(y.cpu().numpy() - ref)) < 1e-12 * max(1.0, np.abs(ref).max())


def test_spmv_row_ranges_and_nonzero_stream_vs_scipy(pb):
    """The products the eigensolver is built from -- a rank's row range, the kernel over the nonzero stream (rows cut
    by tile boundaries, hubs spanning several tiles, empty rows) -- against scipy."""
    import torch
    from scipy.sparse import random as sprandom
    from palantir_b200 import _device as dv, _native

    rng = np.random.default_rng(5)
    n = 30_000
    a = sprandom(n, n, density=50 / n, random_state=5, format="lil")
    for r, ln in ((0, 3000), (5, 129), (77, 1024), (78, 1), (n - 1, 5000), (1234, 2048), (20_000, 0), (20_001, 0)):
        cols = rng.choice(n, ln, replace=False)
        a.rows[r] = sorted(cols.tolist()); a.data[r] = rng.normal(size=ln).tolist()
    a = a.tocsr()
    x = rng.normal(size=n)
    ref = a @ x
    scale = max(1.0, np.abs(ref).max())
    indptr, indices, vals = (dv.to_device(a.indptr.astype(np.int32)), dv.to_device(a.indices.astype(np.int32)),
                             dv.to_device(a.data))
    xd = dv.to_device(x)
    lens = np.diff(a.indptr)
    long_rows = np.flatnonzero(lens > int(_native.load().pb200_spmv_long_row())).astype(np.int32)
    # three row ranges with their own hub lists
    y = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    for r0, r1 in ((0, 9_984), (9_984, 20_000), (20_000, n)):
        mine = long_rows[(long_rows >= r0) & (long_rows < r1)]
        dv.call("pb200_csr_spmv_rows_f64", indptr, indices, vals, n, a.nnz, r0, r1,
                dv.to_device(mine) if len(mine) else None, len(mine), xd, y)
    assert np.max(np.abs(y.cpu().numpy() - ref)) < 1e-12 * scale
    # nonzero stream, whole matrix and the same three row ranges
    tile = int(_native.load().pb200_csr_spmv_stream_tile())
    n_tiles = -(-a.nnz // tile)
    tile_lo = dv.empty((n_tiles + 1,), dv.I32)
    head, tail = dv.zeros((n_tiles,), dv.F64), dv.zeros((n_tiles,), dv.F64)
    dv.call("pb200_csr_spmv_stream_prepare", indptr, n, a.nnz, tile_lo)
    tl = tile_lo.cpu().numpy()
    assert np.array_equal(tl[:-1], np.searchsorted(a.indptr, np.arange(n_tiles) * tile, side="left")) and tl[-1] == n
    y = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    dv.call("pb200_csr_spmv_stream_f64", indptr, indices, vals, n, a.nnz, tile_lo, head, tail, xd, y, 0, n, 0, n_tiles)
    assert np.max(np.abs(y.cpu().numpy() - ref)) < 1e-12 * scale
    y2 = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    for r0, r1 in ((0, 9_984), (9_984, 20_000), (20_000, n)):
        e0, e1 = int(a.indptr[r0]), int(a.indptr[r1])
        dv.call("pb200_csr_spmv_stream_f64", indptr, indices, vals, n, a.nnz, tile_lo, head, tail, xd, y2, r0, r1,
                e0 // tile, -(-e1 // tile))
    assert torch.equal(y, y2)


def test_row_sharded_eigensolver_pieces_reproduce_the_batched_steps(pb):
    """pb200_csr_spmv_rows_f64 / pb200_csr_spmv_stream_f64 + pb200_lanczos_post_pro, step by step (what the row-sharded
    eigensolver runs between its all-gathers), against pb200_lanczos_steps_pro on the same matrix and start vector."""
    import torch
    from palantir_b200 import _device as dv, _native

    z = golden("dm_f32_n1500_d20.npz")
    k = csr_matrix((z["k_data"], z["k_indices"], z["k_indptr"]), shape=(1500, 1500))
    kd = dv.csr_from_host(k)
    n, nnz, steps = 1500, kd.nnz, 120
    tvals, svals, dis, dsq = dv.transition_and_symmetric(kd)
    start = dv.to_device(np.random.default_rng(1).random(n))
    nblk = n // 2048 + 1
    tile = int(_native.load().pb200_csr_spmv_stream_tile())
    n_tiles = -(-nnz // tile)
    tile_lo = dv.empty((n_tiles + 1,), dv.I32)
    head, tail = dv.zeros((n_tiles,), dv.F64), dv.zeros((n_tiles,), dv.F64)
    dv.call("pb200_csr_spmv_stream_prepare", kd.indptr, n, nnz, tile_lo)

    def run(mode):
        V = dv.zeros((steps + 2, n), dv.F64)
        alpha, beta = dv.zeros((steps + 2,), dv.F64), dv.zeros((steps + 3,), dv.F64)
        w, h = dv.empty((n,), dv.F64), dv.empty((steps + 2,), dv.F64)
        partial, tmp = dv.empty(((steps + 2) * nblk,), dv.F64), dv.empty((1,), dv.F64)
        omstride = steps + 3
        om = dv.zeros((3 * omstride,), dv.F64); om[0] = 1.0
        gate, counters = dv.zeros((2,), dv.I32), dv.zeros((1,), dv.I32)
        dv.call("pb200_lanczos_init", start, n, V, None, partial, tmp)
        eps = 2.2e-16 * np.sqrt(n)
        if mode == "batched":
            dv.call("pb200_lanczos_steps_pro", kd.indptr, kd.indices, svals, n, nnz, None, 0, V, n, None, 0, 0, steps,
                    alpha, beta, w, h, partial, tmp, om, omstride, gate, counters, eps, 1.5e-8)
        else:
            for j in range(steps):
                if mode == "rows":
                    for r0, r1 in ((0, 496), (496, 1008), (1008, n)):
                        dv.call("pb200_csr_spmv_rows_f64", kd.indptr, kd.indices, svals, n, nnz, r0, r1, None, 0, V[j], w)
                else:
                    dv.call("pb200_csr_spmv_stream_f64", kd.indptr, kd.indices, svals, n, nnz, tile_lo, head, tail, V[j],
                            w, 0, n, 0, n_tiles)
                dv.call("pb200_lanczos_post_pro", n, V, n, j, alpha, beta, w, h, partial, tmp, om, omstride, gate,
                        counters, eps, 1.5e-8)
        return alpha.cpu().numpy()[:steps], beta.cpu().numpy()[1:steps + 1], V.cpu().numpy(), int(counters.item())

    a0, b0, V0, c0 = run("batched")
    a1, b1, V1, c1 = run("rows")
    assert np.array_equal(a0, a1) and np.array_equal(b0, b1) and np.array_equal(V0, V1) and c0 == c1
    # semi-orthogonal basis, and the tridiagonal matrix carries the leading eigenvalue of T (= 1)
    G = V0[:steps] @ V0[:steps].T
    assert np.abs(G - np.eye(steps)).max() < 1e-6
    from scipy.linalg import eigh_tridiagonal
    assert abs(eigh_tridiagonal(a0, b0[:-1], eigvals_only=True).max() - 1.0) < 1e-10
    # the stream kernel sums a row in another order: same recurrence up to round-off
    a2, b2, V2, c2 = run("stream")
    assert np.abs(a2[:10] - a0[:10]).max() < 1e-12 and np.abs(b2[:10] - b0[:10]).max() < 1e-12
    ev0 = eigh_tridiagonal(a0, b0[:-1], eigvals_only=True)[-3:]
    ev2 = eigh_tridiagonal(a2, b2[:-1], eigvals_only=True)[-3:]
    assert abs(ev2[-1] - 1.0) < 1e-10 and np.abs(ev2 - ev0).max() < 1e-8


def test_windowed_spmv_and_relabelled_matrix_vs_scipy(pb):
    """The eigensolver's SpMV on the relabelled matrix (x window in shared memory) against scipy, on a
    matrix large enough to take the windowed kernel, with columns inside and far outside the window."""
    import torch
    from scipy.sparse import coo_matrix
    from palantir_b200 import _device as dv

    rng = np.random.default_rng(11)
    n, per_row = 150_000, 24
    rows = np.repeat(np.arange(n), per_row)
    near = rows + rng.integers(-9000, 9000, rows.size)
    far = rng.integers(0, n, rows.size)
    cols = np.where(rng.random(rows.size) < 0.8, near, far) % n
    a = coo_matrix((rng.normal(size=rows.size), (rows, cols)), shape=(n, n)).tocsr()
    a.sum_duplicates()
    x = rng.normal(size=n)
    indptr, indices, vals = (dv.to_device(a.indptr.astype(np.int32)), dv.to_device(a.indices.astype(np.int32)),
                             dv.to_device(a.data))
    y = dv.empty((n,), dv.F64)
    dv.call("pb200_csr_spmv_window_f64", indptr, indices, vals, n, a.nnz, dv.to_device(x), y)
    ref = a @ x
    assert np.max(np.abs(y.cpu().numpy() - ref)) < 1e-11 * np.abs(ref).max()
    # relabelling: (P A P^T)(P x) = P (A x)
    perm = rng.permutation(n).astype(np.int32)
    inv = np.empty(n, np.int32); inv[perm] = np.arange(n, dtype=np.int32)
    lens = np.diff(a.indptr)[perm]
    ip = np.zeros(n + 1, np.int32); ip[1:] = np.cumsum(lens)
    ind_p, val_p = dv.empty((a.nnz,), dv.I32), dv.empty((a.nnz,), dv.F64)
    dv.call("pb200_csr_permute", indptr, indices, vals, n, dv.to_device(perm), dv.to_device(inv), dv.to_device(ip),
            ind_p, val_p)
    y2 = dv.empty((n,), dv.F64)
    dv.call("pb200_csr_spmv_window_f64", dv.to_device(ip), ind_p, val_p, n, a.nnz, dv.to_device(x[perm]), y2)
    assert np.max(np.abs(y2.cpu().numpy() - ref[perm])) < 1e-11 * np.abs(ref).max()


def test_magic_imputation_vs_oracle(pb, po):
    z = golden("dm_f64_n600_d8.npz")
    res = pb.utils.run_diffusion_maps(pd.DataFrame(z["x"]), kernel_backend="sklearn")
    rng = np.random.default_rng(0)
    x = rng.random((600, 37)).astype(np.float32)
    got = pb.utils.run_magic_imputation(x, dm_res=res, n_steps=3, sparse=False)
    ref = po.magic_impute(res["T"], x, 3)
    assert got.dtype == np.float32
    near_clip = np.abs(ref - 1e-2) < 1e-5
    assert np.max(np.abs(got - ref)[~near_clip]) <= 1e-5
    sp = pb.utils.run_magic_imputation(pd.DataFrame(x), dm_res=res)
    assert isinstance(sp, pd.DataFrame) and sp.shape == x.shape
    assert isinstance(pb.utils.run_magic_imputation(x, dm_res=res), csr_matrix)
    # float64 X keeps its precision (the reference casts only T^n to float32, utils.py:791)
    x64 = rng.random((600, 301))                            # more than one 128-column tile, ragged tail
    got64 = pb.utils.run_magic_imputation(x64, dm_res=res, n_steps=2, sparse=False)
    ref64 = po.magic_impute(res["T"], x64, 2)
    assert got64.dtype == np.float64 and ref64.dtype == np.float64
    near_clip = np.abs(ref64 - 1e-2) < 1e-5
    assert np.max(np.abs(got64 - ref64)[~near_clip]) <= 1e-5
    # sparse X: densified chunk by chunk, same numbers, reference's return types
    xs = csr_matrix(x * (x > 0.7))
    a = pb.utils.run_magic_imputation(xs, dm_res=res, n_steps=3, sparse=False)
    b = po.magic_impute(res["T"], np.asarray(xs.todense()), 3)
    assert isinstance(a, np.matrix) and np.max(np.abs(np.asarray(a) - b)[np.abs(b - 1e-2) > 1e-5]) <= 1e-5
    assert isinstance(pb.utils.run_magic_imputation(xs, dm_res=res), csr_matrix)


def test_magic_imputation_at_scale_with_chunks(pb, monkeypatch):
    """100 k cells x 700 genes: the operator is relabelled along the diffusion components (locality for the gathers) and
    the genes go through the device in chunks; T (T (T X)) in FP64 on the host is the yardstick (the reference's
    T**3 would hold 2e8 entries here)."""
    from palantir_b200.datasets import synth_branching

    n, g = 100_000, 700
    x, _ = synth_branching(n, 50, seed=0)
    res = pb.utils.run_diffusion_maps(pd.DataFrame(x), kernel_backend="sklearn")
    genes = np.random.default_rng(3).random((n, g), dtype=np.float32)
    monkeypatch.setattr(pb.utils, "MAGIC_CHUNK_GENES", 300)
    got = pb.utils.run_magic_imputation(genes, dm_res=res, n_steps=3, sparse=False)
    t = res["T"]
    cols = np.r_[0:5, 298:303, 695:700]
    ref = t @ (t @ (t @ genes[:, cols].astype(np.float64)))
    assert got.shape == (n, g) and got.dtype == np.float32
    assert np.max(np.abs(got[:, cols] - ref)) <= 1e-5
    # without eigenvectors (operator in input order) the numbers are the same
    got2 = pb.utils.run_magic_imputation(genes[:, :64], dm_res={"T": t}, n_steps=3, sparse=False)
    assert np.max(np.abs(got2 - got[:, :64])) <= 2e-6


# ------------------------------------------------------------------ upstream / downstream of the path
@pytest.mark.parametrize("dtype,sparse_in", [(np.float64, False), (np.float32, False), (np.float64, True)])
def test_run_pca_vs_truncated_svd(pb, po, dtype, sparse_in, monkeypatch):
    """run_pca (utils.py:52-118) against sklearn's TruncatedSVD(arpack), which scanpy's zero_center=False path
    calls (scanpy is not installed here: see oracle.truncated_svd_pca).  Scores / components up to sign."""
    from palantir_b200.compat import AnnDataLike

    rng = np.random.default_rng(4)
    n, g, k = 3000, 180, 25
    lat = rng.normal(size=(n, 12)) * np.linspace(4, 1, 12)
    x = np.maximum(lat @ rng.normal(size=(12, g)) + 0.3 * rng.normal(size=(n, g)) + 1.0, 0).astype(dtype)
    xin = csr_matrix(x) if sparse_in else x
    monkeypatch.setattr(pb.utils, "PCA_ROW_BLOCK", 1024)                 # several row blocks
    ad = AnnDataLike(xin, obs_names=[f"c{i}" for i in range(n)])
    proj, ratio = pb.utils.run_pca(ad, n_components=k, use_hvg=False)
    scores, comps, var, vr = po.truncated_svd_pca(xin, k)
    tol = 1e-8 if dtype == np.float64 else 2e-3
    assert proj.shape == (n, k) and list(proj.index) == list(ad.obs_names) and proj.values.dtype == dtype
    sgn = np.sign(np.sum(proj.values * scores, axis=0))
    assert np.max(np.abs(proj.values * sgn - scores)) <= tol * np.abs(scores).max()
    assert np.max(np.abs(ratio - vr)) <= tol and np.max(np.abs(ad.uns["pca"]["variance"] - var)) <= tol * var.max()
    assert np.max(np.abs(ad.varm["PCs"].T * sgn[:, None] - comps)) <= max(tol, 1e-6)
    if dtype == np.float64 and not sparse_in:
        assert np.array_equal(sgn, np.ones(k))                            # sklearn's svd_flip convention reproduced
    # DataFrame input, other key, highly-variable-gene path with the 85 % cut
    ad2 = AnnDataLike(x, obs_names=ad.obs_names)
    ad2.var["highly_variable"] = np.arange(g) % 3 != 0
    ad2.obsm["X_pca"] = np.zeros((n, 2))
    p2, r2 = pb.utils.run_pca(ad2, n_components=k, use_hvg=True, pca_key="X_hvg")
    assert np.array_equal(ad2.obsm["X_pca"], np.zeros((n, 2))) and ad2.obsm["X_hvg"].shape == p2.shape
    _, _, _, vr_full = po.truncated_svd_pca(x[:, ad2.var["highly_variable"].values], min(1000, n - 1, int(ad2.var["highly_variable"].sum()) - 1))
    want = int(np.where(np.cumsum(vr_full) > 0.85)[0][0])
    assert p2.shape[1] == want and ad2.varm["PCs"].shape == (g, want)
    assert np.all(ad2.varm["PCs"][~ad2.var["highly_variable"].values] == 0)
    p3, r3 = pb.utils.run_pca(pd.DataFrame(x, index=ad.obs_names), n_components=7, use_hvg=False)
    assert p3.shape == (n, 7) and np.max(np.abs(np.abs(p3.values) - np.abs(proj.values[:, :7]))) <= tol * np.abs(scores).max()


@pytest.mark.parametrize("save_as_df", [True, False])
def test_results_survive_the_on_disk_round_trip(pb, names1500, tmp_path, save_as_df):
    """The AnnData keys the path writes (core.py:237-247; north_star) written to disk and read back: with
    SAVE_AS_DF the fate probabilities are a DataFrame whose columns are the terminal cells, without it an array
    plus uns['palantir_fate_probabilities_columns'].  anndata / h5py are not installed in the build image, so the
    stand-in's .npz container with the same key layout is used; with the real anndata the same test goes through
    write_h5ad / read_h5ad."""
    from palantir_b200.compat import AnnDataLike

    z = golden("palantir_n1500_wp200.npz")
    names = names1500
    ad = AnnDataLike(np.zeros((1500, 3)), obs_names=names)
    ad.obsm["DM_EigenVectors_multiscaled"] = z["ms"]
    pb.core.run_palantir(ad, names[int(z["early"])], terminal_states=list(names[z["terminals"]]), num_waypoints=200,
                         save_as_df=save_as_df)
    path = str(tmp_path / "ad.npz")
    ad.write(path)
    back = AnnDataLike.read(path)
    assert np.array_equal(back.obs["palantir_pseudotime"].values, ad.obs["palantir_pseudotime"].values)
    assert np.max(np.abs(back.obs["palantir_pseudotime"].values - z["pseudotime"])) <= 1e-10
    assert np.array_equal(back.obs["palantir_entropy"].values, ad.obs["palantir_entropy"].values)
    assert list(back.uns["palantir_waypoints"]) == list(ad.uns["palantir_waypoints"])
    fp = back.obsm["palantir_fate_probabilities"]
    if save_as_df:
        assert isinstance(fp, pd.DataFrame) and list(fp.columns) == list(names[z["fate_columns"]])
        assert "palantir_fate_probabilities_columns" not in back.uns
        vals = fp.values
    else:
        assert isinstance(fp, np.ndarray)
        assert list(back.uns["palantir_fate_probabilities_columns"]) == list(names[z["fate_columns"]])
        vals = fp
    assert np.max(np.abs(vals - z["fate_probs"])) <= 1e-9
    try:
        import anndata
    except ImportError:
        return
    real = anndata.AnnData(np.zeros((1500, 3)))
    real.obs_names = names
    real.obsm["DM_EigenVectors_multiscaled"] = z["ms"]
    pb.core.run_palantir(real, names[int(z["early"])], terminal_states=list(names[z["terminals"]]), num_waypoints=200,
                         save_as_df=save_as_df)
    real.write_h5ad(str(tmp_path / "ad.h5ad"))
    again = anndata.read_h5ad(str(tmp_path / "ad.h5ad"))
    assert np.max(np.abs(again.obs["palantir_pseudotime"].values - z["pseudotime"])) <= 1e-10
    assert np.max(np.abs(np.asarray(again.obsm["palantir_fate_probabilities"]) - z["fate_probs"])) <= 1e-9


# ------------------------------------------------------------------ benchmark-size properties
def test_properties_at_benchmark_scale(pb):
    """Config 2 (100k cells x 50 PCs, 1200 waypoints): properties that hold at any size."""
    from palantir_b200.datasets import synth_branching, synth_names

    n = 100_000
    x, info = synth_branching(n, 50, seed=0)
    names = synth_names(n)
    res = pb.utils.run_diffusion_maps(pd.DataFrame(x, index=names), kernel_backend="sklearn")
    k, t = res["kernel"], res["T"]
    assert abs(k - k.T).max() == 0.0                                       # K = W + W^T is exactly symmetric
    assert np.all(np.diff(k.indptr) >= 30) and k.has_sorted_indices
    assert np.max(np.abs(np.asarray(t.sum(axis=1)).ravel() - 1.0)) < 1e-12   # T row-stochastic
    lam = res["EigenValues"].values
    assert abs(lam[0] - 1.0) < 1e-10 and np.all(np.diff(lam) <= 0)
    v = res["EigenVectors"].values
    resid = np.linalg.norm(t @ v - v * lam, axis=0)
    assert resid.max() < 1e-8                                              # converged eigenpairs of T
    ms = pb.utils.determine_multiscale_space(res)
    term = {names[p]: b for b, p in info["terminals"].items()}
    pr = pb.core.run_palantir(ms, names[info["early"]], terminal_states=term, num_waypoints=1200)
    pt = pr.pseudotime.values
    assert pt.min() == 0.0 and pt.max() == 1.0
    assert spearmanr(pt, info["time"])[0] > 0.98
    bp = pr.branch_probs.values
    assert bp.shape == (n, 3) and np.all(bp >= 0) and np.all(bp.sum(1) <= 1 + 1e-9) and np.all(bp.sum(1) > 0.97)
    for cell in term:
        assert pr.branch_probs.loc[cell].max() > 0.9
    assert pr.entropy.values.min() >= 0
