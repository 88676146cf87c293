"""Diffusion-map API of Palantir on B200: same signatures, defaults, return types and
AnnData keys as ``palantir.utils`` (reference: src/palantir/utils.py), with the
arithmetic done by the CUDA library through ``_device``.

Functions: compute_kernel (utils.py:340), diffusion_maps_from_kernel (:454),
run_diffusion_maps (:498), determine_multiscale_space (:849),
run_magic_imputation (:727).
"""
from __future__ import annotations

from typing import Dict, Optional, Union
from warnings import warn

import numpy as np
import pandas as pd
import torch
from scipy.sparse import csr_matrix, issparse

from . import _devcache
from . import _device as dv
from . import config as palantir_config
from .compat import is_anndata

__all__ = ["run_pca", "compute_kernel", "diffusion_maps_from_kernel", "run_diffusion_maps",
           "determine_multiscale_space", "run_magic_imputation", "early_cell", "fallback_terminal_cell",
           "find_terminal_states", "CellNotFoundException"]


class CellNotFoundException(Exception):
    """No cell could be determined by the method used (reference utils.py:23)."""

# last-call diagnostics (kNN flagged rows, Lanczos steps ...) for the bench / tests
LAST_INFO: dict = {}


def _fingerprint(m: csr_matrix):
    """Cheap identity of a host CSR matrix: buffer addresses, sizes and a strided sample of the values.
    A cached device copy is used only while this still matches (a matrix edited in place -- rescaled,
    thresholded -- is uploaded again)."""
    d = m.data
    return (m.shape, m.nnz, d.ctypes.data, m.indices.ctypes.data, m.indptr.ctypes.data,
            float(d[::4099].sum()), float(d[:2048].sum()), float(d[-2048:].sum()))


def _remember_device_copy(kernel: csr_matrix, kern: "dv.DeviceCSR") -> None:
    kernel._pb200_device = (kern, _fingerprint(kernel))


def _device_copy_of(kernel) -> Optional["dv.DeviceCSR"]:
    hit = getattr(kernel, "_pb200_device", None) if palantir_config.REUSE_DEVICE_RESULTS else None
    if hit is None:
        return None
    kern, fp = hit
    return kern if fp == _fingerprint(kernel) else None


# ---------------------------------------------------------------------------
# run_pca (reference utils.py:29-118): the step right before the path
# ---------------------------------------------------------------------------
PCA_ROW_BLOCK = 262144      # rows uploaded at a time (a sparse X is densified block by block)


def _truncated_svd(X, n_comps: int):
    """scanpy's ``sc.pp.pca(zero_center=False)`` = sklearn ``TruncatedSVD(algorithm="arpack")``: the leading
    n_comps singular triplets of X (cells x genes, dense or scipy sparse) WITHOUT centring.  X^T X is accumulated
    on the device in FP64 over row blocks, its G x G eigenproblem is solved on the host, the scores X V are
    computed on the device; signs as sklearn's ``svd_flip(u_based_decision=False)`` (largest |entry| of every
    component positive).  Returns (scores [n, k] in X's float type, components [k, g], variance, variance_ratio)."""
    sparse_in = issparse(X)
    if sparse_in:
        X = X.tocsr()
    n, g = X.shape
    ftype = np.float64 if X.dtype == np.float64 else np.float32
    dev_t = dv.F64 if ftype == np.float64 else dv.F32
    C = dv.zeros((g, g), dv.F64)
    colsum = dv.zeros((g,), dv.F64)

    def block(r0):
        r1 = min(n, r0 + PCA_ROW_BLOCK)
        b = X[r0:r1].toarray() if sparse_in else np.asarray(X[r0:r1])
        return dv.to_device(np.ascontiguousarray(b, dtype=ftype))

    blocks = {}
    for r0 in range(0, n, PCA_ROW_BLOCK):
        xb = block(r0)
        dv.call("pb200_pca_gram", xb, int(ftype == np.float64), xb.shape[0], g, g, C)
        colsum += xb.sum(dim=0, dtype=torch.float64)
        if n <= PCA_ROW_BLOCK:
            blocks[r0] = xb                          # a single block stays on the device for the projection
    Ch = C.cpu().numpy()
    Ch = np.triu(Ch) + np.triu(Ch, 1).T            # the kernel fills the tiles on or above the diagonal
    lam, vec = np.linalg.eigh(Ch)
    order = np.argsort(lam)[::-1][:n_comps]
    lam, vec = np.maximum(lam[order], 0.0), vec[:, order]
    comps = vec.T.copy()                            # [k, g]
    signs = np.sign(comps[np.arange(comps.shape[0]), np.argmax(np.abs(comps), axis=1)])
    signs[signs == 0] = 1.0
    comps *= signs[:, None]
    Vd = dv.to_device(np.ascontiguousarray(comps.T))            # [g, k]
    k = comps.shape[0]
    scores = np.empty((n, k), dtype=ftype)
    for r0 in range(0, n, PCA_ROW_BLOCK):
        xb = blocks.get(r0)
        if xb is None:
            xb = block(r0)
        yb = dv.empty((xb.shape[0], k), dv.F64)
        for q0 in range(0, xb.shape[0], 1 << 21):                # grid.y limit of the projection kernel
            q1 = min(xb.shape[0], q0 + (1 << 21))
            dv.call("pb200_pca_project", xb[q0:q1], int(ftype == np.float64), q1 - q0, g, g, Vd, k, yb[q0:q1])
        scores[r0:r0 + xb.shape[0]] = yb.to(dev_t).cpu().numpy()
    mean = colsum.cpu().numpy() / n
    var = lam / n - (comps @ mean) ** 2            # np.var of the scores: E[y^2] - E[y]^2 with y = X v
    full_var = float((np.diag(Ch) / n - mean ** 2).sum())
    return scores, comps, var, var / full_var


def _slice_pca(ad, n_comps: int) -> None:
    """Keep the first n_comps components of stored PCA results (reference utils.py:29-49)."""
    if "X_pca" in ad.obsm:
        ad.obsm["X_pca"] = ad.obsm["X_pca"][:, :n_comps]
    if "pca" in ad.uns:
        for key in ("variance", "variance_ratio"):
            if key in ad.uns["pca"]:
                ad.uns["pca"][key] = np.asarray(ad.uns["pca"][key])[:n_comps]
        params = ad.uns["pca"].get("params")
        if isinstance(params, dict) and "n_comps" in params:
            params["n_comps"] = n_comps
    if "PCs" in getattr(ad, "varm", {}):
        ad.varm["PCs"] = ad.varm["PCs"][:, :n_comps]


def _pca_into(ad, n_comps: int, mask=None) -> None:
    """What ``sc.pp.pca(ad, n_comps, zero_center=False, mask_var=...)`` writes: obsm['X_pca'], varm['PCs'],
    uns['pca'] = {params, variance, variance_ratio}."""
    X = ad.X
    if mask is not None:
        X = X[:, np.asarray(mask, dtype=bool)]
    scores, comps, var, ratio = _truncated_svd(X, n_comps)
    ad.obsm["X_pca"] = scores
    if not hasattr(ad, "varm") or ad.varm is None:
        ad.varm = {}
    if mask is not None:
        pcs = np.zeros((ad.X.shape[1], comps.shape[0]), dtype=comps.dtype)
        pcs[np.asarray(mask, dtype=bool)] = comps.T
    else:
        pcs = comps.T.copy()
    ad.varm["PCs"] = pcs
    ad.uns["pca"] = {"params": {"zero_center": False, "n_comps": int(comps.shape[0])},
                     "variance": var, "variance_ratio": ratio}


def run_pca(data, n_components: int = 300, use_hvg: bool = True, pca_key: str = "X_pca"):
    """PCA of cells x genes data (reference utils.py:52-118).  A DataFrame is wrapped like the reference wraps it
    in an AnnData; with ``use_hvg`` the components come from the genes flagged in ``var['highly_variable']``, 1000
    of them first, cut where the cumulative variance ratio passes 0.85.  Returns (projections DataFrame,
    variance ratio) and, for an AnnData, writes obsm[pca_key], varm['PCs'], uns['pca'] as scanpy does."""
    from .compat import AnnDataLike

    if isinstance(data, pd.DataFrame):
        ad = AnnDataLike(data.values, obs_names=data.index, var_names=data.columns)
        old_pca = None
    else:
        ad = data
        old_pca = ad.obsm["X_pca"].copy() if (pca_key != "X_pca" and "X_pca" in ad.obsm) else None
    n_obs, n_vars = ad.X.shape
    if not use_hvg:
        n_comps = min(n_components, n_obs - 1, n_vars - 1)
        _pca_into(ad, n_comps)
    else:
        var = getattr(ad, "var", None)
        if var is None or "highly_variable" not in var:
            raise KeyError("highly_variable")         # scanpy: mask_var='highly_variable' needs ad.var['highly_variable']
        mask = np.asarray(var["highly_variable"], dtype=bool)
        l_n_comps = min(1000, n_obs - 1, n_vars - 1, int(mask.sum()) - 1)
        _pca_into(ad, l_n_comps, mask)
        try:
            n_comps = np.where(np.cumsum(ad.uns["pca"]["variance_ratio"]) > 0.85)[0][0]
        except IndexError:
            n_comps = n_components
        n_comps = min(n_comps, n_obs - 1, n_vars - 1)
        if n_comps < l_n_comps:
            _slice_pca(ad, n_comps)
        elif n_comps > l_n_comps:
            _pca_into(ad, n_comps, mask)
    pca_projections = pd.DataFrame(ad.obsm["X_pca"], index=ad.obs_names)
    variance_ratio = ad.uns["pca"]["variance_ratio"]
    if is_anndata(data) and pca_key != "X_pca":
        data.obsm[pca_key] = ad.obsm["X_pca"]
        if old_pca is not None:
            data.obsm["X_pca"] = old_pca
        else:
            del data.obsm["X_pca"]
    return pca_projections, variance_ratio


def _knn_table(values: np.ndarray, n_neighbors: int):
    """Full neighbour table on the device, query rows sharded across ranks when a
    process group is up (one process per GPU; allgather of the row shards)."""
    if values.dtype not in (np.float32, np.float64):
        values = values.astype(np.float64)
    x = dv.to_device(values)
    with dv.stage("knn_total"):
        idx, dist, stats = dv.knn_like_sklearn(x, n_neighbors, shard=True)
    order = stats.pop("perm", None)
    LAST_INFO["knn"] = stats
    return idx, dist, order


def _compute_kernel_device(values: np.ndarray, knn: int, alpha: float) -> dv.DeviceCSR:
    n = values.shape[0]
    idx, dist, order = _knn_table(values, min(knn + 1, n))
    kern = dv.adaptive_kernel(idx, dist, knn, alpha)
    kern.order = order
    return kern


def compute_kernel(data, knn: int = 30, alpha: float = 0, pca_key: str = "X_pca",
                   kernel_key: str = "DM_Kernel", backend: Optional[str] = None) -> csr_matrix:
    """Adaptive anisotropic diffusion kernel (reference utils.py:340-451).

    Both backend names run the exact-kNN semantics of the reference's "sklearn"
    branch (utils.py:396-417): k+1 exact neighbours, self stripped, sigma = distance
    to the floor(knn/3)-th neighbour, W = exp(-d/sigma), K = W + W^T, optional
    D^-alpha K D^-alpha.  ("scanpy" in the reference is an approximate search.)"""
    if is_anndata(data):
        data_df = data.obsm[pca_key]
        if not isinstance(data_df, pd.DataFrame):
            data_df = pd.DataFrame(data_df, index=data.obs_names)
    else:
        data_df = data
    backend = backend or palantir_config.KERNEL_BACKEND
    if backend not in {"scanpy", "sklearn"}:
        raise ValueError("backend must be one of {'scanpy', 'sklearn'}")
    if issparse(data_df):
        raise ValueError("compute_kernel on B200 needs dense PCA projections (sparse input is not supported)")
    if knn is None or int(knn) <= 0:
        raise ValueError("knn must be a positive integer")
    values = data_df.values if isinstance(data_df, pd.DataFrame) else np.asarray(data_df)
    kern = _compute_kernel_device(values, int(knn), float(alpha))
    kernel = dv.csr_to_host(kern)
    _remember_device_copy(kernel, kern)  # lets diffusion_maps_from_kernel skip the re-upload
    if is_anndata(data):
        data.obsp[kernel_key] = kernel
    return kernel


def _diffusion_maps_device(kern: dv.DeviceCSR, n_components: int, seed, t_home: Optional[list] = None):
    """(T values, eigenvalues, eigenvectors) on the device.  With ``t_home`` (a list) the copy of T
    to the host is started before the eigensolver and its finisher appended to the list."""
    n = kern.n
    symmetric = kern.symmetric or dv.check_symmetric(kern)
    tvals, svals, dis, dsq = dv.transition_and_symmetric(kern)
    # same RNG use as the reference (utils.py:482-483)
    np.random.seed(seed)
    v0 = np.random.rand(n)
    if t_home is not None:
        t_home.append(dv.csr_to_host_async(kern, tvals))
    if symmetric:
        # symmetric Lanczos on S = D^-1/2 K D^-1/2; the start vector is mapped to the symmetric problem:
        # K_m(T, v0) = D^-1/2 K_m(S, D^1/2 v0)
        start = dv.empty((n,), dv.F64)
        dv.call("pb200_vec_mul", dv.to_device(v0), dsq, start, n)
        with dv.stage("eigensolver"):
            lam, U, info = dv.lanczos_topk(kern, svals, dis, start, n_components)
    else:
        # a kernel that is not symmetric (the reference takes any matrix, utils.py:477-484): Arnoldi on T itself
        with dv.stage("eigensolver"):
            lam, U, info = dv.arnoldi_topk(kern, tvals, dv.to_device(v0), n_components)
    from . import _dist

    if _dist.world()[1] > 1:
        # every rank returns rank 0's eigenpairs bit for bit (the ranks go on to shard one job between them)
        lam_t = _dist.broadcast_(torch.from_numpy(np.ascontiguousarray(lam, dtype=np.float64)).to(U.device))
        lam = lam_t.cpu().numpy()
        U = _dist.broadcast_(U.contiguous())
    LAST_INFO["lanczos"] = info
    return tvals, lam, U


def _eigenvector_frame(U: torch.Tensor) -> pd.DataFrame:
    """Host DataFrame of the eigenvectors; the device original is remembered for determine_multiscale_space."""
    host = dv.to_host(U)
    _devcache.remember(host, U)
    return pd.DataFrame(host, copy=False)


def diffusion_maps_from_kernel(kernel: csr_matrix, n_components: int = 10,
                               seed: Union[int, None] = 0) -> Dict[str, object]:
    """Diffusion map of a kernel matrix (reference utils.py:454-495): returns
    ``{"T", "EigenVectors", "EigenValues"}`` with T = D^-1 K (csr), the top
    eigenvectors of T as unit-L2 columns (DataFrame) and the eigenvalues, descending
    (Series).  Eigenvectors are converged (the reference stops ARPACK at tol 1e-4);
    their signs are arbitrary, as ARPACK's are."""
    kern = _device_copy_of(kernel)
    if kern is None:
        kern = dv.csr_from_host(kernel)
    tvals, lam, U = _diffusion_maps_device(kern, n_components, seed)
    T = dv.csr_to_host(kern, tvals)
    return {"T": T, "EigenVectors": _eigenvector_frame(U), "EigenValues": pd.Series(lam)}


def run_diffusion_maps(data, n_components: int = 10, knn: int = 30, alpha: float = 0,
                       seed: Union[int, None] = 0, kernel_backend: str = "scanpy", pca_key: str = "X_pca",
                       kernel_key: str = "DM_Kernel", sim_key: str = "DM_Similarity",
                       eigval_key: str = "DM_EigenValues", eigvec_key: str = "DM_EigenVectors"):
    """Reference utils.py:498-585: kernel + diffusion map; writes obsp[kernel_key],
    obsp[sim_key], obsm[eigvec_key], uns[eigval_key] when given an AnnData and returns
    the dict ``{T, EigenVectors, EigenValues, kernel}`` in every case."""
    if is_anndata(data):
        data_df = pd.DataFrame(data.obsm[pca_key], index=data.obs_names)
    else:
        data_df = data
    if not isinstance(data_df, pd.DataFrame) and not issparse(data_df):
        raise ValueError("'data_df' should be a pd.DataFrame or AnnData")
    if not issparse(data_df):
        # same results as compute_kernel + diffusion_maps_from_kernel, with the two sparse matrices
        # travelling to the host while the eigensolver runs
        if kernel_backend not in {"scanpy", "sklearn"}:
            raise ValueError("backend must be one of {'scanpy', 'sklearn'}")
        if knn is None or int(knn) <= 0:
            raise ValueError("knn must be a positive integer")
        from . import _dist

        kern = _compute_kernel_device(data_df.values, int(knn), float(alpha))
        want_sparse = palantir_config.HOST_SPARSE_ON_ALL_RANKS or _dist.world()[0] == 0
        k_home = dv.csr_to_host_async(kern) if want_sparse else None
        t_home: Optional[list] = [] if want_sparse else None
        tvals, lam, U = _diffusion_maps_device(kern, n_components, seed, t_home)
        vecs = _eigenvector_frame(U)
        kernel = k_home() if want_sparse else None
        if kernel is not None:
            _remember_device_copy(kernel, kern)
        res = {"T": t_home[0]() if want_sparse else None, "EigenVectors": vecs, "EigenValues": pd.Series(lam)}
    else:
        warn("'data' is a sparse matrix and will be interpreted as kernel. "
             "To avoid this warning compute diffusion maps from a precompued kernel using "
             "palantir.utils.diffusion_maps_from_kernel().")
        kernel = data_df
        res = diffusion_maps_from_kernel(kernel, n_components, seed)
    res["kernel"] = kernel
    if not issparse(data_df):
        res["EigenVectors"].index = data_df.index
    if is_anndata(data):
        data.obsp[kernel_key] = res["kernel"]
        data.obsp[sim_key] = res["T"]
        data.obsm[eigvec_key] = res["EigenVectors"].values
        data.uns[eigval_key] = res["EigenValues"].values
    return res


def determine_multiscale_space(dm_res, n_eigs: Union[int, None] = None, eigval_key: str = "DM_EigenValues",
                               eigvec_key: str = "DM_EigenVectors",
                               out_key: str = "DM_EigenVectors_multiscaled") -> pd.DataFrame:
    """Reference utils.py:849-911: eigen-gap choice of n_eigs (>= 3) and the rescaling
    V[:, 1:n_eigs] * lam/(1-lam).  An N x 10 array times a handful of scalars: host
    arithmetic, as in the reference (0.004 s at 100k cells).  Returns the DataFrame in
    both the dict and the AnnData case (reference behaviour; its docstring says None)."""
    if is_anndata(dm_res):
        vecs = dm_res.obsm[eigvec_key]
        if not isinstance(vecs, pd.DataFrame):
            vecs = pd.DataFrame(vecs, index=dm_res.obs_names)
        d = {"EigenValues": dm_res.uns[eigval_key], "EigenVectors": vecs}
    else:
        d = dm_res
    if not isinstance(d, dict):
        raise ValueError("'dm_res' should be a dict or a AnnData instance")
    if n_eigs is None:
        vals = np.ravel(d["EigenValues"])
        gaps = vals[:-1] - vals[1:]
        n_eigs = np.argsort(gaps)[-1] + 1
        if n_eigs < 3:
            n_eigs = np.argsort(gaps)[-2] + 1
        if n_eigs < 3:
            n_eigs = 3
    n_eigs = int(n_eigs)
    ev = np.ravel(np.asarray(d["EigenValues"]))[1:n_eigs]
    vecs = d["EigenVectors"].values
    u_dev = _devcache.lookup(vecs)
    if u_dev is not None and u_dev.shape == vecs.shape:
        # the eigenvectors are still on the device (and unchanged on the host): same IEEE products there
        scale = torch.from_numpy(np.ascontiguousarray(ev / (1 - ev))).to(u_dev.device)
        ms_dev = (u_dev[:, 1:n_eigs] * scale).contiguous()
        ms_host = dv.to_host(ms_dev)
        _devcache.remember(ms_host, ms_dev)
        out = pd.DataFrame(ms_host, index=d["EigenVectors"].index, copy=False)
    else:
        # columns 1..n_eigs-1 (utils.py:905-908); a slice, so the product is the only pass over the data
        out = pd.DataFrame(vecs[:, 1:n_eigs] * (ev / (1 - ev)), index=d["EigenVectors"].index, copy=False)
    if is_anndata(dm_res):
        dm_res.obsm[out_key] = out.values
    return out


MAGIC_CHUNK_GENES = 4096        # gene columns resident on the device at a time (2 x N x chunk x 4 B)
MAGIC_ORDER_MIN_CELLS = 65536   # from this size on the operator is relabelled into a locality-preserving order


def _magic_operator(T, eigenvectors=None):
    """Device copy of the diffusion operator for MAGIC: CSR with FP32 values (the reference casts T^n to float32,
    utils.py:791) and -- at scale, when diffusion components are at hand -- relabelled along the Morton curve of the
    leading components, so that the X rows a neighbourhood of output rows gathers stay in L2.
    Returns (indptr, indices, vals32, n, perm or None)."""
    from . import _device_core as dc

    op = dv.csr_from_host(T)
    n = op.n
    perm = None
    indptr, indices, vals = op.indptr, op.indices, op.data
    if eigenvectors is not None and n >= MAGIC_ORDER_MIN_CELLS and eigenvectors.shape[1] >= 2:
        comps = np.ascontiguousarray(np.asarray(eigenvectors)[:, 1:4], dtype=np.float64)
        perm, inv, _ = dc.morton_order(dv.to_device(comps))
        lens = (op.indptr[1:] - op.indptr[:-1])[perm.long()]
        indptr = torch.zeros((n + 1,), dtype=dv.I32, device=perm.device)
        indptr[1:] = torch.cumsum(lens, 0)
        indices = dv.empty((op.nnz,), dv.I32)
        vals = dv.empty((op.nnz,), dv.F64)
        dv.call("pb200_csr_permute", op.indptr, op.indices, op.data, n, perm, inv, indptr, indices, vals)
    vals32 = dv.empty((op.nnz,), dv.F32)
    dv.call("pb200_cast_f64_f32", vals, vals32, op.nnz)
    return indptr, indices, vals32, n, perm


def magic_device(operator, x: torch.Tensor, g: int, n_steps: int, clip_threshold: Optional[float] = None) -> torch.Tensor:
    """T^n_steps X as n_steps successive gene-tiled SpMMs on a device matrix x [n, ld] (FP32 or FP64, ld a multiple
    of 4, first g columns used; rows in the operator's order).  Returns the result buffer (same shape)."""
    indptr, indices, vals32, n, _ = operator
    ld = x.shape[1]
    cur, nxt = x, torch.empty_like(x)
    if ld > g:
        nxt[:, g:] = 0
    is64 = int(x.dtype == torch.float64)
    with dv.stage("magic_spmm"):
        for _ in range(int(n_steps)):
            dv.call("pb200_csr_spmm_tiled", indptr, indices, vals32, n, cur, ld, nxt, ld, g, is64)
            cur, nxt = nxt, cur
        if clip_threshold is not None:
            dv.call("pb200_clip_below", cur, cur.numel(), float(clip_threshold), is64)
    return cur


def run_magic_imputation(data, dm_res: Union[dict, None] = None, n_steps: int = 3, sim_key: str = "DM_Similarity",
                         expression_key: str = None, imputation_key: str = "MAGIC_imputed_data",
                         n_jobs: int = -1, sparse: bool = True, clip_threshold: float = 1e-2):
    """MAGIC imputation T^n_steps X (reference utils.py:727-846) as n_steps successive SpMMs with the
    device-resident operator -- T^n is never materialised (the reference builds it: 1978 nnz/row at n=3).
    The genes are processed in column chunks (a sparse X is densified one chunk at a time, as the reference's
    100-column chunks do); X keeps its precision (float32 stays float32, float64 stays float64: the reference
    casts only T^n to float32); with several ranks every rank imputes its own share of the columns.  Return type
    follows the input type as in the reference; values below clip_threshold are zeroed."""
    from . import _dist

    if is_anndata(data):
        if expression_key is not None:
            if expression_key not in data.layers.keys():
                raise ValueError(f"expression_key '{expression_key}' not found in .layers.")
            X = data.layers[expression_key]
        else:
            X = data.X
        T = data.obsp[sim_key] if dm_res is None else None
    elif isinstance(data, pd.DataFrame):
        X = data.values
        T = None
    else:
        X = data
        T = None
    vecs = None
    if dm_res is not None:
        T = dm_res["T"]
        ev = dm_res.get("EigenVectors") if isinstance(dm_res, dict) else None
        if ev is not None:
            vecs = ev.values if isinstance(ev, pd.DataFrame) else np.asarray(ev)
    elif not is_anndata(data):
        raise ValueError("Diffusion map results (dm_res) must be provided if data is not AnnData")
    elif "DM_EigenVectors" in getattr(data, "obsm", {}):
        vecs = np.asarray(data.obsm["DM_EigenVectors"])

    x_sparse = issparse(X)
    if x_sparse:
        X = X.tocsc()
    n, g = X.shape
    out_dtype = np.float64 if X.dtype == np.float64 else np.float32
    if vecs is not None and vecs.shape[0] != n:
        vecs = None
    operator = _magic_operator(T, vecs)
    perm = operator[4]
    perm_l = perm.long() if perm is not None else None
    rank, world = _dist.world()
    col_bounds = _dist.shard_bounds(g, world)
    c_lo, c_hi = col_bounds[rank], col_bounds[rank + 1]
    pieces = []
    for c0 in range(c_lo, c_hi, MAGIC_CHUNK_GENES):
        c1 = min(c_hi, c0 + MAGIC_CHUNK_GENES)
        w = c1 - c0
        ld = dv.round_up(w, 4)
        chunk = X[:, c0:c1].toarray() if x_sparse else np.asarray(X[:, c0:c1])
        xh = torch.from_numpy(np.ascontiguousarray(chunk, dtype=out_dtype))
        xd = dv.zeros((n, ld), dv.F64 if out_dtype == np.float64 else dv.F32) if ld != w else \
            dv.empty((n, ld), dv.F64 if out_dtype == np.float64 else dv.F32)
        xd[:, :w].copy_(xh, non_blocking=False)
        if perm_l is not None:
            xd = xd[perm_l].contiguous()
        yd = magic_device(operator, xd, w, n_steps, clip_threshold)
        if perm_l is not None:
            back = torch.empty_like(yd)
            back[perm_l] = yd
            yd = back
        pieces.append(yd[:, :w])
        del xd
    mine = torch.cat(pieces, dim=1) if len(pieces) != 1 else pieces[0]
    if world > 1:
        wmax = max(col_bounds[r + 1] - col_bounds[r] for r in range(world))
        pad = dv.zeros((n, wmax), mine.dtype)
        pad[:, :mine.shape[1]] = mine
        full = dv.empty((world, n, wmax), mine.dtype)
        torch.distributed.all_gather_into_tensor(full, pad)
        mine = torch.cat([full[r, :, :col_bounds[r + 1] - col_bounds[r]] for r in range(world)], dim=1)
        if not (palantir_config.HOST_SPARSE_ON_ALL_RANKS or rank == 0):
            return None
    imputed = dv.to_host(mine.contiguous())

    if sparse:
        imputed = csr_matrix(imputed)                   # values below the threshold are already exact zeros
    elif x_sparse:
        imputed = np.asmatrix(imputed)
    if is_anndata(data):
        data.layers[imputation_key] = imputed if issparse(imputed) else np.asarray(imputed)
    if isinstance(data, pd.DataFrame):
        dense_out = imputed.todense() if issparse(imputed) else imputed
        imputed = pd.DataFrame(dense_out, index=data.index, columns=data.columns)
    return imputed


# ---------------------------------------------------------------------------
# start / terminal cell helpers (reference utils.py:914-1137): thin callers of the path
# ---------------------------------------------------------------------------
def _column_extremes(eigenvectors: np.ndarray):
    """(argmax, argmin) position of every column, first occurrence like numpy's argmax / argmin;
    one device pass over the matrix (its device original is used when it is still there)."""
    dev = _devcache.lookup(eigenvectors)
    if dev is None or dev.dtype != torch.float64 or dev.shape != eigenvectors.shape:
        vals = eigenvectors if eigenvectors.dtype == np.float64 else eigenvectors.astype(np.float64)
        dev = dv.to_device(vals)
    from . import _device_core as dc

    _, imin, _, imax = dc.col_minmax(dev)
    return imax.cpu().numpy(), imin.cpu().numpy()


def early_cell(ad, celltype: str, celltype_column: str = "celltype",
               eigvec_key: str = "DM_EigenVectors_multiscaled", fallback_seed: Optional[int] = None):
    """The cell of ``celltype`` sitting at an extreme of a diffusion component (reference
    utils.py:944-1031): components are tried in order, maximum before minimum; with
    ``fallback_seed`` the reverse-pseudotime search of ``fallback_terminal_cell`` is used when no
    component qualifies, otherwise ``CellNotFoundException`` is raised."""
    if not is_anndata(ad):
        raise ValueError("'ad' should be an instance of AnnData")
    if eigvec_key not in ad.obsm:
        raise ValueError(
            f"'{eigvec_key}' not found in ad.obsm. "
            "Run `palantir.utils.run_diffusion_maps(ad)` to "
            "compute diffusion map eigenvectors.")
    eigenvectors = ad.obsm[eigvec_key]
    if isinstance(eigenvectors, pd.DataFrame):
        eigenvectors = eigenvectors.values
    if not isinstance(celltype_column, str):
        raise ValueError(f"celltype_column='{celltype_column}' should be a string")
    if celltype_column not in ad.obs.columns:
        raise ValueError(f"celltype_column='{celltype_column}' should be a column of ad.obs.")
    if not isinstance(celltype, str):
        raise ValueError("celltype should be a string")
    labels = ad.obs[celltype_column]
    if celltype not in labels.values:
        raise ValueError(f"Celltype '{celltype}' not found in ad.obs['{celltype_column}'].")
    if fallback_seed is not None and not isinstance(fallback_seed, int):
        raise ValueError("'fallback_seed' should be an integer")

    imax, imin = _column_extremes(np.asarray(eigenvectors))
    for dcomp in range(eigenvectors.shape[1]):
        for pos, which in ((int(imax[dcomp]), "max"), (int(imin[dcomp]), "min")):
            if labels.iloc[pos] == celltype:
                cell = ad.obs_names[pos]
                print(f"Using {cell} for cell type {celltype} which is {which} in "
                      f"diffusion component {dcomp}.")
                return cell
    if fallback_seed is not None:
        print("Falling back to slow early cell detection.")
        return fallback_terminal_cell(ad, celltype, celltype_column=celltype_column, eigvec_key=eigvec_key,
                                      seed=fallback_seed)
    raise CellNotFoundException(
        f"No valid component found: {celltype} "
        "Consider increasing the number of diffusion components "
        "('n_components' in palantir.utils.run_diffusion_maps) "
        "or specify a 'fallback_seed' to determine an early cell based on "
        f"reverse pseudotime starting from random non-{celltype} cell.")


def fallback_terminal_cell(ad, celltype: str, celltype_column: str = "anno",
                           eigvec_key: str = "DM_EigenVectors_multiscaled", seed: int = 2353):
    """Reference utils.py:1034-1083: start Palantir from a random cell of another type (pandas'
    ``sample(1, random_state=seed)``) and return the cell of ``celltype`` with the largest pseudotime.
    Like the reference call, this writes the run's results into ``ad``."""
    from .core import run_palantir

    other_cells = ad.obs_names[ad.obs[celltype_column] != celltype]
    fake_early_cell = other_cells.to_series().sample(1, random_state=seed).iloc[0]
    pr_res = run_palantir(ad, fake_early_cell, eigvec_key=eigvec_key, terminal_states=None,
                          use_early_cell_as_start=True)
    idx = ad.obs[celltype_column] == celltype
    ec = pr_res.pseudotime[idx].argmax()
    cell = ad.obs_names[idx][ec]
    print(f"Using {cell} for cell type {celltype} which is latest cell in "
          f"{celltype} when starting from {fake_early_cell}.")
    return cell


def find_terminal_states(ad, celltypes, celltype_column: str = "celltype",
                         eigvec_key: str = "DM_EigenVectors_multiscaled", fallback_seed: Optional[int] = None):
    """Reference utils.py:1086-1137: ``early_cell`` for every cell type; types without a qualifying
    component are skipped with a warning.  Returns a Series cell -> cell type."""
    terminal_states = pd.Series(dtype=str)
    for ct in celltypes:
        try:
            cell = early_cell(ad, ct, celltype_column, eigvec_key, fallback_seed)
        except CellNotFoundException:
            warn(f"No valid component found: {ct} "
                 "Consider increasing the number of diffusion components "
                 "('n_components' in palantir.utils.run_diffusion_maps). "
                 f"The cell type {ct} will be skipped.")
            continue
        terminal_states[cell] = ct
    return terminal_states
