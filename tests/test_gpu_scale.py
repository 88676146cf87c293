"""GPU parity at the benchmark's own sizes (BASELINE.json configs 2 and 3) and across ranks.

At 1 M cells the whole reference cannot be run in a test, so config 3 is checked stage by stage on
samples the CPU finishes in seconds (SURVEY.md 8d "P-stage"): sampled kNN rows against sklearn's brute
force over all references, the kernel against the oracle's construction from the same neighbour table,
the eigenpairs through their residual on the host copy of T, and shortest paths from three sources
against scipy's Dijkstra on the same graph.  Config 2 (100 k cells) is run end to end against the
converged ("tight") oracle.  The two-rank test needs two GPUs and is skipped otherwise.
"""
import contextlib
import io
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest
from scipy.linalg import subspace_angles
from scipy.sparse import csr_matrix
from scipy.stats import spearmanr

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


@pytest.fixture(scope="module")
def c3(pb):
    """Config 3 input (1 M cells x 100 PCs) and one device-resident run of the whole path."""
    from palantir_b200 import _device as dv
    from palantir_b200 import pipeline
    from palantir_b200.datasets import synth_branching

    n, d = 1_000_000, 100
    x, info = synth_branching(n, d, seed=0)
    x_dev = dv.to_device(x)
    out = pipeline.run_device(x_dev, info["early"], list(info["terminals"].values()))
    return dict(x=x, info=info, x_dev=x_dev, out=out, n=n, d=d)


def test_c3_knn_sampled_rows_vs_sklearn_brute(pb, c3):
    """2,000 sampled query rows of the 1 M x 100 search (utils.py:404-407): same neighbour sets as
    sklearn's brute force over all 1 M references, distances bit-equal (float32 input rule)."""
    from sklearn.neighbors import NearestNeighbors

    from palantir_b200 import _device as dv

    x, n = c3["x"], c3["n"]
    idx, dist, stats = dv.knn_like_sklearn(c3["x_dev"], 31)
    assert stats["flagged"] == 0 or stats["flagged"] < n // 1000
    rows = np.random.default_rng(5).choice(n, 2000, replace=False)
    rows_d = dv.to_device(rows.astype(np.int64))
    got_i = idx[rows_d].cpu().numpy()
    got_d = dist[rows_d].cpu().numpy()
    rd, ri = NearestNeighbors(n_neighbors=31, algorithm="brute").fit(x).kneighbors(x[rows])
    # sklearn's chunked DGEMM sums x.y in an order that depends on the chunk shapes, so a 2,000-row query block
    # may differ from the full-table run (which the kernel reproduces bit for bit, see the fixtures) in the
    # last bit of d^2 -- visible after the float32 rounding of the distance only for a handful of entries, and
    # in the self distance (pure cancellation noise around 1e-8)
    nz = rd[:, 1:] > 0
    rel = np.abs(got_d[:, 1:] - rd[:, 1:])[nz] / rd[:, 1:][nz]
    assert rel.max() <= 2.4e-7                                   # two float32 ulps
    assert np.mean(got_d[:, 1:] != rd[:, 1:]) < 1e-3
    assert np.max(np.abs(got_d[:, 0] - rd[:, 0])) < 1e-5
    same = (np.sort(got_i, 1) == np.sort(ri, 1)).all(1)
    near_ties = np.array([np.min(np.diff(r[1:])) <= 2.4e-7 * r[-1] for r in rd])
    assert np.all(same | near_ties) and (~same).sum() < 20
    c3["knn"] = (idx, dist)


def test_c3_kernel_vs_oracle_from_same_neighbours(pb, po, c3):
    """Same neighbour table on both sides (utils.py:408-417, 440-446): identical pattern, values <= 1e-12 rel."""
    from palantir_b200 import _device as dv

    if "knn" not in c3:
        c3["knn"] = dv.knn_like_sklearn(c3["x_dev"], 31)[:2]
    idx, dist = c3["knn"]
    kern = dv.adaptive_kernel(idx, dist, 30, 0.0)
    got = dv.csr_to_host(kern)
    want = po.kernel_from_knn(dist.cpu().numpy(), idx.cpu().numpy().astype(np.int64), 30)
    want.sort_indices()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.max(np.abs(got.data - want.data) / want.data) <= 1e-12
    # and it is the kernel the pipeline run produced
    assert np.array_equal(c3["out"]["kernel"].indices.cpu().numpy(), got.indices)


def test_c3_eigenpairs_residual(pb, c3):
    """T v = lambda v on the host copy of T (utils.py:478-495): residual far below the 1e-5 eigenvalue bar."""
    from palantir_b200 import _device as dv

    out = c3["out"]
    t = dv.csr_to_host(out["kernel"], out["t_values"])
    v = out["eigenvectors"].cpu().numpy()
    lam = np.asarray(out["eigenvalues"])
    assert np.all(np.diff(lam) <= 0) and abs(lam[0] - 1.0) < 1e-9
    assert np.allclose(np.linalg.norm(v, axis=0), 1.0, atol=1e-12)
    resid = np.linalg.norm(t @ v - v * lam, axis=0)
    assert resid.max() < 1e-8
    rs = np.asarray(t.sum(axis=1)).ravel()
    assert np.max(np.abs(rs - 1.0)) < 1e-12


def test_c3_shortest_paths_vs_scipy_dijkstra(pb, c3):
    """Three sources on the 1 M-cell multiscale kNN graph (core.py:355-362) against csgraph.dijkstra on the
    same graph: <= 1e-12 relative (the cell-decomposed solver adds path segments in another order than a heap
    Dijkstra does; the per-source kernel is bit-equal)."""
    from scipy.sparse.csgraph import dijkstra

    from palantir_b200 import _device as dv
    from palantir_b200 import _device_core as dc

    out, n = c3["out"], c3["n"]
    dev = dc.minmax_scale(out["multiscale"])
    perm, inv, work = dc.morton_order(dev)
    idx, dist, _ = dc.knn_table_boxed(work, 30)
    g = dc.connect_graph(idx, dist, work, int(inv[c3["info"]["early"]].item()))
    src = np.array([0, n // 3, n - 7], dtype=np.int64)
    D, _ = dc.sssp(g, dv.to_device(src))
    host = csr_matrix((g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy()), shape=(n, n))
    want = dijkstra(host, directed=False, indices=src)
    got = D.cpu().numpy()
    assert np.all(np.isfinite(got) == np.isfinite(want))
    fin = np.isfinite(want)
    assert np.max(np.abs(got[fin] - want[fin]) / np.maximum(want[fin], 1e-300)) <= 1e-12
    # the many-source path (what run_palantir uses at this size) agrees with the few-source one
    src2 = np.concatenate([src, np.random.default_rng(1).choice(n, 637, replace=False)])
    D2, _ = dc.sssp(g, dv.to_device(src2))
    got2 = D2[:3].cpu().numpy()
    assert np.max(np.abs(got2[fin] - want[fin]) / np.maximum(want[fin], 1e-300)) <= 1e-12


def test_c3_results_are_sane(pb, c3):
    out, info, n = c3["out"], c3["info"], c3["n"]
    pt = out["pseudotime"].cpu().numpy()
    fp = out["fate_probs"].cpu().numpy()
    assert pt.min() == 0.0 and pt.max() == 1.0
    assert spearmanr(pt[::5], info["time"][::5])[0] > 0.98
    assert fp.shape == (n, 3) and np.all(fp >= 0) and np.all(np.abs(fp.sum(1) - 1) < 1e-9)
    for p in info["terminals"].values():
        assert fp[p].max() > 0.9


def test_c1_sized_full_run_vs_oracles(pb, po):
    """Config 1's size (4142 cells x 100 PCs, 500 waypoints: the reference's tutorial data set, here the synthetic
    generator since the bundled h5ad is not in the checkout) through the public API: against the converged oracle
    at the north_star tolerances, and against the reference AS SHIPPED (ARPACK tol 1e-4) for the eigenvalues, as
    the reference's own test does (tests/test_utils_diffusion_maps_from_kernel.py:44-52, atol 1e-4)."""
    from palantir_b200.datasets import synth_branching, synth_names

    n, d, nwp = 4142, 100, 500
    x, info = synth_branching(n, d, seed=0)
    names = synth_names(n)
    term = list(info["terminals"].values())
    with contextlib.redirect_stdout(io.StringIO()):
        res = pb.utils.run_diffusion_maps(pd.DataFrame(x, index=names), n_components=10, knn=30,
                                          kernel_backend="sklearn")
        ms = pb.utils.determine_multiscale_space(res)
        pr = pb.core.run_palantir(ms, names[info["early"]], terminal_states=list(names[term]), num_waypoints=nwp)
        o = po.run_diffusion_maps(x, 10, 30, tol=1e-12)
        asis = po.run_diffusion_maps(x, 10, 30)
        oms = po.multiscale_space(o["EigenVectors"], o["EigenValues"])
        op = po.run_palantir(oms, info["early"], terminal=term, num_waypoints=nwp)
    assert np.array_equal(res["kernel"].indices, o["kernel"].indices)
    assert np.max(np.abs(res["kernel"].data - o["kernel"].data) / o["kernel"].data) <= 1e-12
    ev = res["EigenValues"].values
    assert np.max(np.abs(ev - o["EigenValues"]) / np.abs(o["EigenValues"])) <= 1e-5
    assert np.allclose(ev, asis["EigenValues"], atol=1e-4)
    assert subspace_angles(res["EigenVectors"].values, o["EigenVectors"]).max() < 1e-4
    assert ms.shape == oms.shape
    assert list(names.get_indexer(pr.waypoints)) == list(op["waypoints"])
    assert np.max(np.abs(pr.pseudotime.values - op["pseudotime"])) <= 1e-4
    assert spearmanr(pr.pseudotime.values, op["pseudotime"])[0] > 0.999
    assert np.max(np.abs(pr.branch_probs.values - op["fate_probs"])) <= 1e-4


def test_c2_full_run_vs_tight_oracle(pb, po):
    """Config 2 (100 k cells x 50 PCs, 1200 waypoints) through the public API against the converged oracle
    (ARPACK tol 1e-12; the reference as shipped stops at 1e-4 and is itself ~1e-3 rad from its converged
    self): eigenvalues 1e-5 rel, subspace angle < 1e-4, waypoints identical, pseudotime and
    fate probabilities <= 1e-4 (north_star tolerances)."""
    from palantir_b200.datasets import synth_branching, synth_names

    n, d, nwp = 100_000, 50, 1200
    x, info = synth_branching(n, d, seed=0)
    names = synth_names(n)
    term = list(info["terminals"].values())
    with contextlib.redirect_stdout(io.StringIO()):
        res = pb.utils.run_diffusion_maps(pd.DataFrame(x, index=names), n_components=10, knn=30,
                                          kernel_backend="sklearn")
        ms = pb.utils.determine_multiscale_space(res)
        pr = pb.core.run_palantir(ms, names[info["early"]], terminal_states=list(names[term]), num_waypoints=nwp)
        o = po.run_diffusion_maps(x, 10, 30, tol=1e-12)
        oms = po.multiscale_space(o["EigenVectors"], o["EigenValues"])
        op = po.run_palantir(oms, info["early"], terminal=term, num_waypoints=nwp, row_apply_entropy=False)
    assert np.array_equal(res["kernel"].indices, o["kernel"].indices)
    assert np.max(np.abs(res["kernel"].data - o["kernel"].data) / o["kernel"].data) <= 1e-12
    ev = res["EigenValues"].values
    assert np.max(np.abs(ev - o["EigenValues"]) / np.abs(o["EigenValues"])) <= 1e-5
    assert subspace_angles(res["EigenVectors"].values, o["EigenVectors"]).max() < 1e-4
    assert ms.shape == oms.shape
    assert subspace_angles(ms.values, oms).max() < 1e-4
    assert list(names.get_indexer(pr.waypoints)) == list(op["waypoints"])
    assert np.max(np.abs(pr.pseudotime.values - op["pseudotime"])) <= 1e-4
    assert spearmanr(pr.pseudotime.values, op["pseudotime"])[0] > 0.999
    assert list(names.get_indexer(pr.branch_probs.columns)) == list(op["terminal"])
    assert np.max(np.abs(pr.branch_probs.values - op["fate_probs"])) <= 1e-4
    assert np.max(np.abs(pr.entropy.values - op["entropy"])) <= 1e-4


def test_two_ranks_reproduce_single_gpu(pb, tmp_path):
    """The sharded run (torchrun, 2 ranks, NCCL: kNN rows and shortest-path sources split) returns the
    single-GPU results: kernel and waypoints identical, pseudotime / fates within 1e-9."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ref = str(tmp_path / "ref.npz")
    script = os.path.join(ROOT, "scripts", "mgpu_check.py")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0,1"))
    r = subprocess.run([sys.executable, script, "--save", ref], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", script, "--compare", ref],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
