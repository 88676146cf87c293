"""CPU: host-side mirror (validation, config, PResults, multiscale space, sharding, ABI exports)."""
import os
import re
import subprocess
import sys
import warnings

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT, golden


def test_library_exports_every_declared_symbol():
    from palantir_b200 import _native

    decl = _native.parse_header()
    assert len(decl) >= 40
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (pb200_\w+)", out))
    missing = sorted(set(decl) - exported)
    assert not missing, missing
    lib = _native.load()                      # loads and binds every prototype (no compute call)
    assert lib.pb200_abi_version() == 1


def test_no_cpu_fallback_without_cuda():
    import torch

    from palantir_b200 import _native, utils

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_native.NativeLibraryError):
        utils.compute_kernel(pd.DataFrame(np.random.rand(40, 20)), backend="sklearn")


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "palantir_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_multiscale_space_matches_oracle_and_reference_rules():
    from oracle import palantir_oracle as po
    from palantir_b200 import utils

    z = golden("dm_f32_n1500_d20.npz")
    res = {"EigenValues": pd.Series(z["evals_asis"]), "EigenVectors": pd.DataFrame(z["evecs_asis"])}
    ms = utils.determine_multiscale_space(res)
    assert isinstance(ms, pd.DataFrame)
    assert np.array_equal(ms.values, z["ms_asis"])
    assert utils.determine_multiscale_space(res, n_eigs=3).shape[1] == 2      # reference test :20-22
    with pytest.raises(ValueError):
        utils.determine_multiscale_space("not a dict")
    assert po.pick_n_eigs(z["evals_asis"]) - 1 == ms.shape[1]


def test_multiscale_space_anndata_like():
    from palantir_b200 import utils
    from palantir_b200.compat import AnnDataLike, is_anndata

    z = golden("dm_f32_n1500_d20.npz")
    ad = AnnDataLike(np.zeros((1500, 3)))
    assert is_anndata(ad) and not is_anndata(pd.DataFrame()) and not is_anndata({})
    ad.obsm["DM_EigenVectors"] = z["evecs_asis"]
    ad.uns["DM_EigenValues"] = z["evals_asis"]
    out = utils.determine_multiscale_space(ad)
    assert np.array_equal(ad.obsm["DM_EigenVectors_multiscaled"], out.values)
    assert list(out.index) == list(ad.obs_names)


def test_input_validation_errors():
    from palantir_b200 import core, utils

    with pytest.raises(ValueError):
        utils.run_diffusion_maps("nope")
    with pytest.raises(ValueError):
        utils.compute_kernel(pd.DataFrame(np.random.rand(10, 3)), backend="faiss")
    with pytest.raises(ValueError):
        utils.run_magic_imputation(np.zeros((3, 3)))
    df = pd.DataFrame(np.random.rand(20, 4), index=[f"c{i}" for i in range(20)])
    with pytest.raises(ValueError):
        core.run_palantir(df, "c0", knn=0)
    with pytest.raises(ValueError):
        core.run_palantir(df, "missing_cell")


def test_normalize_cell_identifiers():
    from palantir_b200 import config
    from palantir_b200.validation import normalize_cell_identifiers as norm

    obs = pd.Index(["1", "2", "3"])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        d, nreq, nfound = norm([1, 2, 9], obs)
    assert set(d) == {"1", "2", "9"} and (nreq, nfound) == (3, 2) and len(w) >= 1
    d, nreq, nfound = norm(pd.Series({"1": "A"}), obs)
    assert d == {"1": "A"} and nfound == 1
    d, _, nfound = norm(["1"], pd.Index([1, 2, 3]))
    assert nfound == 1                              # obs_names treated as strings
    config.AUTO_CONVERT_CELL_IDS_TO_STR = False
    try:
        with warnings.catch_warnings(record=True):
            warnings.simplefilter("always")
            _, _, nfound = norm([1], obs)
        assert nfound == 0
    finally:
        config.AUTO_CONVERT_CELL_IDS_TO_STR = True
    with pytest.raises(ValueError):
        norm(3.5, obs)
    with pytest.raises(ValueError):
        norm([], obs)


def test_presults_clips_small_probabilities(tmp_path):
    from palantir_b200.presults import PResults

    bp = pd.DataFrame([[0.005, 0.995], [0.5, 0.5]])
    pr = PResults(pd.Series([0.0, 1.0]), pd.Series([0.1, 0.7]), bp, pd.Index([0]))
    assert pr.branch_probs.iloc[0, 0] == 0 and pr.branch_probs.iloc[0, 1] == 0.995
    f = tmp_path / "r.pkl"
    pr.save(str(f))
    assert PResults.load(str(f)).branch_probs.equals(pr.branch_probs)


def test_sklearn_auto_rule():
    from palantir_b200._device import sklearn_uses_brute

    assert sklearn_uses_brute(1000, 16, 31) and not sklearn_uses_brute(1000, 15, 31)
    assert sklearn_uses_brute(50, 10, 31) and not sklearn_uses_brute(63, 10, 30)


def test_shard_bounds():
    from palantir_b200._dist import shard_bounds

    b = shard_bounds(1_000_000, 8, 256)
    assert b[0] == 0 and b[-1] == 1_000_000 and all(x % 256 == 0 for x in b[1:-1])
    assert all(b[i] <= b[i + 1] for i in range(8))
    assert shard_bounds(100, 4, 256) == [0, 0, 0, 0, 100]
    assert shard_bounds(1185, 8) == [0, 148, 296, 444, 592, 740, 888, 1036, 1185]


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from palantir_b200 import _dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
n, k = 1000, 5
q0, nq = _dist.row_shard(n, 256)
full_i = torch.arange(n * k, dtype=torch.int32).reshape(n, k)
full_d = torch.arange(n * k, dtype=torch.float64).reshape(n, k) * 0.5
gi, gd = _dist.allgather_rows(full_i[q0:q0 + nq].clone(), full_d[q0:q0 + nq].clone(), n, 256)
assert torch.equal(gi, full_i) and torch.equal(gd, full_d), "allgather mismatch"
v = torch.full((7,), float(dist.get_rank() + 1), dtype=torch.float64)
_dist.allreduce_sum_(v)
assert torch.all(v == 3.0)
b = torch.full((3,), float(dist.get_rank()), dtype=torch.float64)
_dist.broadcast_(b, 0)
assert torch.all(b == 0.0)
# waypoint (source) shards cover all rows exactly once
bounds = _dist.shard_bounds(1185, 2)
assert bounds == [0, 592, 1185]
# row shards cut at equal work (uneven, explicit boundaries), one of them empty on the way
from palantir_b200._device import cuts_at_equal_work
import numpy as np
for cuts in (cuts_at_equal_work(np.r_[np.full(4, 10.0), np.full(4, 1.0)], 2, n), [0, 0, n], [0, 768, n]):
    r = dist.get_rank()
    a, b_ = cuts[r], cuts[r + 1]
    gi, gd = _dist.allgather_rows(full_i[a:b_].clone(), full_d[a:b_].clone(), n, 256, cuts)
    assert torch.equal(gi, full_i) and torch.equal(gd, full_d), ("allgather with explicit bounds", cuts)
# the row-sharded eigensolver's exchange: every rank fills its slice of w, one in-place all-gather completes it
chunk = 512
w = torch.full((2 * chunk,), -1.0, dtype=torch.float64)
w[dist.get_rank() * chunk:(dist.get_rank() + 1) * chunk] = torch.arange(chunk, dtype=torch.float64) + 1000.0 * dist.get_rank()
dist.all_gather_into_tensor(w, w[dist.get_rank() * chunk:(dist.get_rank() + 1) * chunk])
assert torch.equal(w, torch.cat([torch.arange(chunk, dtype=torch.float64), torch.arange(chunk, dtype=torch.float64) + 1000.0]))
dist.destroy_process_group()
print("ok", dist.get_rank() if False else sys.argv[3])
'''


def test_sharding_world_size_2_gloo(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = str(s.getsockname()[1])
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err
        assert "ok" in out


def test_early_cell_argument_checks_match_the_reference():
    """The ValueErrors of utils.py:979-1011 are raised before any device work."""
    import pandas as pd
    import pytest
    from palantir_b200 import utils
    from palantir_b200.compat import AnnDataLike

    ad = AnnDataLike(X=np.zeros((4, 1)))
    ad.obs["celltype"] = ["a", "a", "b", "b"]
    with pytest.raises(ValueError, match="should be an instance of AnnData"):
        utils.early_cell(pd.DataFrame(np.zeros((4, 2))), "a")
    with pytest.raises(ValueError, match="not found in ad.obsm"):
        utils.early_cell(ad, "a")
    ad.obsm["DM_EigenVectors_multiscaled"] = np.zeros((4, 2))
    with pytest.raises(ValueError, match="should be a column of ad.obs"):
        utils.early_cell(ad, "a", celltype_column="nope")
    with pytest.raises(ValueError, match="should be a string"):
        utils.early_cell(ad, "a", celltype_column=3)
    with pytest.raises(ValueError, match="celltype should be a string"):
        utils.early_cell(ad, 1)
    with pytest.raises(ValueError, match="not found in ad.obs"):
        utils.early_cell(ad, "c")
    with pytest.raises(ValueError, match="'fallback_seed' should be an integer"):
        utils.early_cell(ad, "a", fallback_seed=1.5)


def test_device_copy_cache_semantics(monkeypatch):
    """palantir_b200/_devcache.py on CPU tensors: off unless config.REUSE_DEVICE_RESULTS; when on, hit by identity
    and by a view of the same memory, miss after an in-place edit, entries die with their host array."""
    import gc

    import torch
    from palantir_b200 import _devcache as dc
    from palantir_b200 import config

    dc.clear()
    off = np.ones((4, 4))
    dc.remember(off, torch.ones(4, 4))
    assert config.REUSE_DEVICE_RESULTS is False and dc.lookup(off) is None and not dc._ENTRIES
    monkeypatch.setattr(config, "REUSE_DEVICE_RESULTS", True)
    a = np.random.default_rng(0).random((5000, 7))
    t = torch.from_numpy(a.copy())
    dc.remember(a, t)
    assert dc.lookup(a) is t
    df = pd.DataFrame(a, copy=False)
    assert dc.lookup(df.values) is t                      # DataFrame.values: a new view object over the same memory
    assert dc.lookup(a.copy()) is None                    # equal numbers, other memory
    assert dc.lookup(a[:, :3]) is None                    # other shape / strides
    a[:, 2] *= -1.0                                        # a sign flip of one column must invalidate the copy
    assert dc.lookup(a) is None
    a[:, 2] *= -1.0
    assert dc.lookup(a) is t                               # restored bit for bit: valid again
    a[17, 3] += 1.0                                        # a single-element edit is caught when it is in the sample ...
    b = np.zeros((3, 3)); dc.remember(b, torch.zeros(3, 3))
    b[1, 1] = 5.0
    assert dc.lookup(b) is None                            # ... and always for small arrays (every element is sampled)
    n_before = len(dc._ENTRIES)
    del a, df, b
    gc.collect()
    assert len(dc._ENTRIES) < n_before
    dc.clear()


def _warning_texts(fn):
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = fn()
    return out, [str(x.message) for x in w]


def test_normalize_cell_identifiers_reference_scenarios():
    """The scenarios of the reference's tests/test_validation.py:250-437, one by one."""
    from palantir_b200 import config
    from palantir_b200.validation import normalize_cell_identifiers as norm

    obs_str = pd.Index(["cell_A", "cell_B", "cell_C", "cell_D", "cell_E"])
    obs_int = pd.Index([100, 200, 300, 400, 500])
    (d, nreq, nfound), msgs = _warning_texts(lambda: norm(["cell_A", "cell_C", "cell_E"], obs_str, context="test"))
    assert (nreq, nfound) == (3, 3) and set(d) == {"cell_A", "cell_C", "cell_E"} and set(d.values()) == {""} and not msgs
    (d, nreq, nfound), msgs = _warning_texts(lambda: norm({"cell_A": "type1", "cell_B": "type2"}, obs_str, context="test"))
    assert d == {"cell_A": "type1", "cell_B": "type2"} and (nreq, nfound) == (2, 2) and not msgs
    (d, nreq, nfound), _ = _warning_texts(lambda: norm(pd.Series(["l1", "l2"], index=["cell_A", "cell_C"]), obs_str))
    assert d == {"cell_A": "l1", "cell_C": "l2"} and (nreq, nfound) == (2, 2)
    # integer ids against string names: converted to str, nothing found, the "none found" warning names both types
    (d, nreq, nfound), msgs = _warning_texts(lambda: norm(pd.Series(["pDC", "ERP"], index=[123, 456]), obs_str, context="test"))
    assert (nreq, nfound) == (2, 0) and "123" in d and "456" in d
    assert any("None of the 2 requested cells were found" in m for m in msgs)
    # string ids against integer names: names compared as strings, with the mismatch warning
    (d, nreq, nfound), msgs = _warning_texts(lambda: norm(["100", "300", "500"], obs_int, context="test"))
    assert (nreq, nfound) == (3, 3) and set(d) == {"100", "300", "500"}
    assert any("Cell identifiers are strings but data.obs_names contains integers" in m for m in msgs)
    (d, nreq, nfound), msgs = _warning_texts(lambda: norm([100, 300, 500], obs_int, context="test"))
    assert (nreq, nfound) == (3, 3) and set(d) == {100, 300, 500} and not msgs
    (d, nreq, nfound), msgs = _warning_texts(lambda: norm(["cell_A", "cell_X", "cell_Y"], obs_str[:3], context="test"))
    assert (nreq, nfound) == (3, 1) and any("Only 1/3 requested cells were found" in m for m in msgs)
    (d, nreq, nfound), msgs = _warning_texts(lambda: norm(["cell_X", "cell_Y", "cell_Z"], obs_str[:3], context="test"))
    assert (nreq, nfound) == (3, 0) and any("None of the 3 requested cells were found" in m for m in msgs)
    config.WARN_ON_CELL_ID_CONVERSION = False
    try:
        (d, nreq, nfound), msgs = _warning_texts(lambda: norm(["cell_X", "cell_Y"], obs_str[:2], context="test"))
        assert (nreq, nfound) == (2, 0) and not msgs
    finally:
        config.WARN_ON_CELL_ID_CONVERSION = True
    config.AUTO_CONVERT_CELL_IDS_TO_STR = False
    try:
        (d, nreq, nfound), msgs = _warning_texts(lambda: norm([100, 200], pd.Index(["100", "200", "300"]), context="test"))
        assert (nreq, nfound) == (2, 0) and any("Automatic conversion is disabled" in m for m in msgs)
    finally:
        config.AUTO_CONVERT_CELL_IDS_TO_STR = True
    for cells in (pd.Index(["cell_A", "cell_C"]), np.array(["cell_A", "cell_C"])):
        (d, nreq, nfound), _ = _warning_texts(lambda: norm(cells, obs_str[:3], context="test"))
        assert (nreq, nfound) == (2, 2) and set(d) == {"cell_A", "cell_C"}
    with pytest.raises(ValueError, match="No cells provided"):
        norm([], obs_str)
    with pytest.raises(ValueError, match="Invalid cells format"):
        norm(12345, obs_str)


def test_normalize_cell_identifiers_differential_vs_reference_source():
    """Where the reference checkout is present (this container, not the GPU box): same return values and the same
    warning texts as src/palantir/validation.py on a grid of inputs."""
    sys.path.insert(0, ROOT)
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference checkout not present")
    ref_shim.load_reference()
    ref_norm = sys.modules[ref_shim._PKG + ".validation"].normalize_cell_identifiers
    from palantir_b200.validation import normalize_cell_identifiers as norm

    obs_sets = [pd.Index(["a", "b", "c"]), pd.Index([1, 2, 3]), pd.Index(["1", "2", "3"])]
    cell_sets = [["a", "c"], ["a", "zz"], [1, 3], [1, 9], ["1", "2"], {"a": "x", 2: "y"}, pd.Series({"b": "B"}),
                 pd.Index(["c"]), np.array([3, 2]), pd.Series(["u", "v"], index=[1, 7])]
    for obs in obs_sets:
        for cells in cell_sets:
            got, gw = _warning_texts(lambda: norm(cells, obs, context="ctx"))
            want, ww = _warning_texts(lambda: ref_norm(cells, obs, context="ctx"))
            assert got[0] == want[0] and got[1:] == want[1:], (list(obs), cells)
            assert gw == ww, (list(obs), cells, gw, ww)


def test_determine_multiscale_space_gap_rules_vs_reference_source():
    """Eigen-gap rules of utils.py:890-903 (clear gap, no clear gap -> second largest, floor of 3) on the eigenvalue
    patterns of the reference's tests/test_utils_determine_multiscale_space.py; compared with the reference source
    itself where the checkout is present."""
    from palantir_b200 import utils

    rng = np.random.default_rng(4)
    clear = np.zeros(10); clear[:4] = [0.95, 0.85, 0.75, 0.30]; clear[4:] = np.linspace(0.25, 0.1, 6)
    flat = np.linspace(0.9, 0.5, 5)
    first = np.array([0.99, 0.5, 0.45, 0.4, 0.39])              # largest gap after the first value -> second largest
    cases = [(clear, 2), (flat, 3), (first, 2)]                  # columns = n_eigs - 1
    sys.path.insert(0, ROOT)
    from oracle import ref_shim

    ref_fn = None
    if ref_shim.available():
        _, ref_utils = ref_shim.load_reference()
        ref_fn = ref_utils.determine_multiscale_space
    for ev, want_cols in cases:
        vecs = pd.DataFrame(rng.random((40, len(ev))), index=[f"c{i}" for i in range(40)])
        res = {"EigenValues": pd.Series(ev), "EigenVectors": vecs}
        out = utils.determine_multiscale_space(res)
        assert out.shape == (40, want_cols) and list(out.index) == list(vecs.index)
        for n_eigs in (None, 3, 4):
            mine = utils.determine_multiscale_space(res, n_eigs=n_eigs)
            if ref_fn is not None:
                theirs = ref_fn(res, n_eigs=n_eigs)
                assert np.array_equal(mine.values, theirs.values) and list(mine.index) == list(theirs.index)


def test_every_abi_entry_point_is_documented():
    """INTEGRATION.md maps every entry point of include/palantir_b200.h to the reference call site it replaces
    (or lists it as infrastructure): the header and the document must not drift apart."""
    header = open(os.path.join(ROOT, "include", "palantir_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 60
    missing = [n for n in names if n not in doc and n.rstrip("12") not in doc]
    assert not missing, missing


def test_install_aliases_the_package_as_palantir():
    """the implementation is importable as `palantir` (the reference itself is not installed here)."""
    import importlib
    import subprocess

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import palantir_b200\n"
        "m = palantir_b200.install()\n"
        "import palantir\n"
        "from palantir.utils import run_diffusion_maps, determine_multiscale_space, run_magic_imputation, run_pca\n"
        "from palantir.core import run_palantir\n"
        "from palantir.presults import PResults\n"
        "import palantir.config\n"
        "assert palantir is palantir_b200 and run_palantir is palantir_b200.core.run_palantir\n"
        "assert palantir.run_palantir is run_palantir and palantir.config.SAVE_AS_DF is True\n"
        "print('ok')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_install_patches_an_installed_reference(monkeypatch):
    """With the reference importable, install() re-binds exactly the functions of the path and leaves the rest."""
    import types

    import palantir_b200

    ref = types.ModuleType("palantir")
    ref.utils, ref.core = types.ModuleType("palantir.utils"), types.ModuleType("palantir.core")
    ref.presults, ref.validation = types.ModuleType("palantir.presults"), types.ModuleType("palantir.validation")
    ref.plot = types.ModuleType("palantir.plot")
    sentinel = object()
    ref.utils.run_diffusion_maps = ref.utils.run_density = sentinel
    ref.core.run_palantir = ref.run_palantir = sentinel
    ref.plot.plot_palantir_results = sentinel
    for name, mod in (("palantir", ref), ("palantir.utils", ref.utils), ("palantir.core", ref.core),
                      ("palantir.presults", ref.presults), ("palantir.validation", ref.validation)):
        monkeypatch.setitem(sys.modules, name, mod)
    out = palantir_b200.install()
    assert out is ref and ref.__pb200_installed__
    assert ref.utils.run_diffusion_maps is palantir_b200.utils.run_diffusion_maps
    assert ref.core.run_palantir is palantir_b200.core.run_palantir and ref.run_palantir is palantir_b200.core.run_palantir
    assert ref.utils.run_density is sentinel and ref.plot.plot_palantir_results is sentinel      # off the path: untouched
    assert ref.presults.PResults is palantir_b200.PResults


def test_knn_shard_cuts_at_equal_work():
    """Row boundaries of the sharded kNN from per-tile work: aligned, monotone, complete, balanced."""
    from palantir_b200._device import cuts_at_equal_work

    rng = np.random.default_rng(0)
    n = 1_000_000
    n_t = (n + 127) // 128
    # work density as along a branching trajectory: between 5 % and 25 % of the tiles visited
    work = 400 + 1500 * (0.5 + 0.5 * np.sin(np.linspace(0, 7, n_t))) + rng.integers(0, 50, n_t)
    for parts in (2, 3, 4, 8):
        cuts = cuts_at_equal_work(work, parts, n)
        assert len(cuts) == parts + 1 and cuts[0] == 0 and cuts[-1] == n
        assert all(a <= b for a, b in zip(cuts, cuts[1:]))
        assert all(c % 256 == 0 for c in cuts[1:-1])
        share = [work[a // 128:(b + 127) // 128].sum() for a, b in zip(cuts, cuts[1:])]
        assert max(share) / (work.sum() / parts) < 1.01
        even = [work[a // 128:(b + 127) // 128].sum() for a, b in zip(range(0, n, n // parts), range(n // parts, n + 1, n // parts))]
        assert max(even) / (work.sum() / parts) > 1.05          # what equal row counts would have given
    # degenerate inputs: all the work in one tile, fewer tile pairs than parts
    w = np.zeros(100); w[7] = 1.0
    cuts = cuts_at_equal_work(w, 4, 100 * 128 - 5)
    assert cuts[0] == 0 and cuts[-1] == 100 * 128 - 5 and all(a <= b for a, b in zip(cuts, cuts[1:]))
    cuts = cuts_at_equal_work(np.ones(3), 8, 300)
    assert cuts[0] == 0 and cuts[-1] == 300 and all(a <= b for a, b in zip(cuts, cuts[1:])) and max(cuts) == 300


def test_bench_traffic_figures_point_at_committed_profiles():
    """Every DRAM-traffic figure bench.py reports names the ncu summary it was read from; the file must be in the tree."""
    import ast

    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    table = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and any(getattr(t, "id", None) == "NCU_TRAFFIC_C3" for t in node.targets):
            table = ast.literal_eval(node.value)
    assert table and {"knn_candidates_tc_kernel", "csr_spmv_stream_kernel", "sssp_cells_stage"} <= set(table)
    for kernel, (nbytes, source) in table.items():
        assert nbytes > 0 and os.path.exists(os.path.join(ROOT, source)), (kernel, source)


def test_stage_ranges_are_inert_by_default_and_nest_with_nvtx(monkeypatch):
    """dv.stage(): no events and no NVTX calls unless asked for; with PB200_NVTX ranges are pushed and popped in pairs."""
    import torch
    from palantir_b200 import _device as dv

    calls = []
    monkeypatch.setattr(torch.cuda.nvtx, "range_push", lambda name: calls.append(("push", name)))
    monkeypatch.setattr(torch.cuda.nvtx, "range_pop", lambda: calls.append(("pop",)))
    monkeypatch.setattr(dv, "PROFILE", None)
    monkeypatch.setattr(dv, "NVTX_RANGES", False)
    with dv.stage("outer"):
        pass
    assert calls == []
    monkeypatch.setattr(dv, "NVTX_RANGES", True)
    with dv.stage("outer"):
        with dv.stage("inner"):
            pass
    assert calls == [("push", "outer"), ("push", "inner"), ("pop",), ("pop",)]
