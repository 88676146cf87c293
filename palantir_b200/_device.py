"""Device-side orchestration of the diffusion-map stages (kNN, kernel, eigensolver).

Everything here works on CUDA tensors and calls the C-ABI kernels through
``_native.call``.  torch owns the buffers and streams and does the O(N)-or-smaller glue
between kernels (prefix sums of row lengths, index gathers, the 100-element dot products of
the principal-axis iteration, concatenations): plumbing, each a few microseconds; every
stage that shows up in the step time is a kernel of the library.  Host objects (scipy /
pandas) are assembled by ``utils.py``.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _native
from ._native import call

I32, I64, F32, F64, U8 = torch.int32, torch.int64, torch.float32, torch.float64, torch.uint8

# timings of the last run, filled when PROFILE is a dict (bench.py sets it)
PROFILE: Optional[dict] = None


def _dev():
    return _native.require_cuda()


def empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=_dev())


def zeros(shape, dtype):
    return torch.zeros(shape, dtype=dtype, device=_dev())


_UPLOAD_MIN_BYTES = 32 << 20      # from here on a pageable host array goes up through the staged pipeline
_UPLOAD_CHUNK = 8 << 20         # 16 staging threads x 8 MB: 13.2 ms for 400 MB (4 x 16 MB: 19.6 ms; scripts/prof_upload.py)
_UPLOAD_BUFS = 16
_upload_state: dict = {}


def _upload_staged(flat: np.ndarray, dst: torch.Tensor) -> None:
    """Pageable host bytes -> device through a ring of page-locked staging buffers: a few host threads copy the
    next chunks into the ring (numpy releases the GIL) while the DMA engine drains the previous ones.  A plain
    cudaMemcpy from pageable memory stages through ONE driver thread (~10 GB/s: 40 ms for the 400 MB PCA matrix of
    config 3, 8 % of the end-to-end step)."""
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor

    dev = dst.device
    st = _upload_state.get(dev.index)
    if st is None:
        # one staging thread per buffer; the ranks of a multi-GPU run on one host share its cores
        from . import _dist
        n_bufs = max(2, min(_UPLOAD_BUFS, (os.cpu_count() or 4) // max(1, _dist.world()[1])))
        st = _upload_state[dev.index] = {
            "n": n_bufs,
            "pool": ThreadPoolExecutor(n_bufs),
            "bufs": [torch.empty(_UPLOAD_CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(n_bufs)],
            "events": [torch.cuda.Event() for _ in range(n_bufs)],
            "stream": torch.cuda.Stream(device=dev)}
    n_bufs = st["n"]
    bufs_np = [b.numpy() for b in st["bufs"]]
    side = st["stream"]
    side.wait_stream(torch.cuda.current_stream(dev))
    nbytes = flat.nbytes
    dst_b = dst.view(torch.uint8).reshape(-1)
    n_chunks = (nbytes + _UPLOAD_CHUNK - 1) // _UPLOAD_CHUNK

    def stage(i):
        k = i % n_bufs
        st["events"][k].synchronize()                  # the DMA that last used this buffer has finished
        a = i * _UPLOAD_CHUNK
        b = min(nbytes, a + _UPLOAD_CHUNK)
        np.copyto(bufs_np[k][: b - a], flat[a:b])
        return k, a, b

    pending = deque()
    for i in range(n_chunks + 1):
        if i < n_chunks:
            pending.append(st["pool"].submit(stage, i))
        while pending and (len(pending) >= n_bufs or i >= n_chunks):
            k, a, b = pending.popleft().result()
            with torch.cuda.stream(side):
                dst_b[a:b].copy_(st["bufs"][k][: b - a], non_blocking=True)
                st["events"][k].record(side)
    torch.cuda.current_stream(dev).wait_stream(side)


def to_device(a: np.ndarray) -> torch.Tensor:
    """Host array -> contiguous row-major device tensor.  A Fortran-ordered matrix (what ``DataFrame.values``
    returns for a frame that owns its data) is uploaded as it lies in memory and transposed on the device instead
    of being re-laid-out by the host first; large arrays go through the staged pipeline of _upload_staged."""
    a = np.asarray(a)
    transposed = a.ndim == 2 and not a.flags.c_contiguous and a.flags.f_contiguous
    src = a.T if transposed else np.ascontiguousarray(a)
    if src.nbytes >= _UPLOAD_MIN_BYTES and src.dtype.kind in "fiu" and src.dtype.itemsize in (1, 2, 4, 8):
        t = torch.empty(src.shape, dtype=torch.from_numpy(src[:0]).dtype, device=_dev())
        _upload_staged(src.reshape(-1).view(np.uint8), t)
    else:
        t = torch.from_numpy(src).to(_dev(), non_blocking=False)
    return t.t().contiguous() if transposed else t


_PIN_MIN_BYTES = 1 << 20


def to_host(*tensors: torch.Tensor):
    """Device tensors -> numpy arrays.  Large results go through page-locked buffers (DMA at PCIe
    speed instead of a staged copy into freshly faulted pageable memory); the arrays returned own
    their pinned buffers.  One synchronisation for the whole batch."""
    outs, pending = [], False
    for t in tensors:
        t = t.contiguous()
        if t.is_cuda and t.numel() * t.element_size() >= _PIN_MIN_BYTES:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t, non_blocking=True)
            pending = True
            outs.append(buf)
        else:
            outs.append(t.cpu())
    if pending:
        torch.cuda.current_stream().synchronize()
    arrs = [o.numpy() for o in outs]
    return arrs[0] if len(arrs) == 1 else arrs


class HostCopy:
    """Device -> host copies in flight on a side stream (see to_host_async)."""

    def __init__(self, bufs, keep, stream):
        self._bufs, self._keep, self._stream = bufs, keep, stream

    def result(self):
        """Wait for the copies; numpy arrays owning their page-locked buffers."""
        if self._stream is not None:
            self._stream.synchronize()
        self._keep = None
        arrs = [b.numpy() for b in self._bufs]
        return arrs[0] if len(arrs) == 1 else arrs


_COPY_STREAMS: dict = {}


def to_host_async(*tensors: torch.Tensor) -> HostCopy:
    """Start copying device tensors into page-locked host buffers on a side stream, ordered after
    the work already queued on the current stream, and return at once: the DMA runs under whatever
    is launched next (the kernel matrix goes home while the eigensolver runs).  ``.result()`` waits."""
    tensors = [t.contiguous() for t in tensors]
    if not tensors or not tensors[0].is_cuda:
        return HostCopy([t.cpu() for t in tensors], None, None)
    dev = tensors[0].device
    side = _COPY_STREAMS.get(dev.index)
    if side is None:
        side = _COPY_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    bufs = []
    with torch.cuda.stream(side):
        for t in tensors:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t, non_blocking=True)
            t.record_stream(side)
            bufs.append(buf)
    return HostCopy(bufs, tensors, side)


def csr_to_host_async(c: "DeviceCSR", data: Optional[torch.Tensor] = None):
    """csr_to_host with the transfer left in flight; call the returned function for the scipy matrix."""
    from scipy.sparse import csr_matrix

    h = to_host_async(c.data if data is None else data, c.indices, c.indptr)
    n = c.n

    def finish():
        vals, indices, indptr = h.result()
        return csr_matrix((vals, indices, indptr), shape=(n, n))

    return finish


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# PB200_NVTX=1: every stage is also an NVTX range (names as in bench.py's stages_ms), so that `ncu --nvtx --nvtx-include
# "knn_candidates/"` or an nsys timeline can address a stage by name
NVTX_RANGES = os.environ.get("PB200_NVTX", "0") == "1"


class _Timer:
    """CUDA-event stage timer (only active when PROFILE is a dict) and, with PB200_NVTX=1, an NVTX range."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX_RANGES:
            torch.cuda.nvtx.range_push(self.name)
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.setdefault("_events", []).append((self.name, self.e0, self.e1))
        if NVTX_RANGES:
            torch.cuda.nvtx.range_pop()
        return False


def stage(name):
    return _Timer(name)


def flush_profile():
    """Resolve recorded events into PROFILE[name] = milliseconds (summed)."""
    if PROFILE is None:
        return
    torch.cuda.synchronize()
    for name, e0, e1 in PROFILE.pop("_events", []):
        PROFILE[name] = PROFILE.get(name, 0.0) + e0.elapsed_time(e1)


# ---------------------------------------------------------------------------
# kNN
# ---------------------------------------------------------------------------
KNN_MAX_KEEP = 48
# candidate-generation engine: "simt" (FP32 FFMA kernel) or "tc" (tcgen05 3xTF32 kernel, d <= 128)
KNN_ENGINE = os.environ.get("PB200_KNN_ENGINE", "tc")


def sklearn_uses_brute(n: int, d: int, n_neighbors: int) -> bool:
    """The 'auto' rule of sklearn NearestNeighbors (sklearn/neighbors/_base.py:618-626):
    brute force (FP64 expanded-form distances) iff d > 15 or n_neighbors >= n // 2,
    otherwise a kd-tree (direct-difference distances)."""
    return d > 15 or n_neighbors >= n // 2


@dataclass
class KnnPrep:
    x: torch.Tensor          # input data [n, d] f32/f64
    perm: Optional[torch.Tensor]   # work order: work point i = input row perm[i] (None = identity)
    xt: torch.Tensor         # centred FP32 copy in work order, tiled k-major (d_pad * ld floats)
    yn: torch.Tensor
    xcn: torch.Tensor
    xn64: torch.Tensor
    rmax2: torch.Tensor
    box_lo: Optional[torch.Tensor]    # [ntiles, 3] projection boxes of the 128-point tiles
    box_hi: Optional[torch.Tensor]
    xtc: Optional[torch.Tensor]       # TF32 hi/lo operand copy for the tensor-core engine
    mean: torch.Tensor
    n: int
    d: int
    d_pad: int
    ld: int


def principal_axes(x: torch.Tensor, mean: torch.Tensor, n_axes: int = 3, iters: int = 5) -> torch.Tensor:
    """Leading principal axes of the centred data ([n_axes, d], orthonormal, norm scaled just below 1)
    by power iteration on X^T X with deflation (two passes over X per step, no host sync).  Only the
    *quality* of the pruning depends on them: any orthonormal set gives valid bounds."""
    from . import _dist

    if x.shape[0] > 250_000:
        # the axes only steer the pruning: a strided sample of ~125k rows pins them well enough
        x = x[:: x.shape[0] // 125_000].contiguous()
    n, d = x.shape
    is64 = int(x.dtype == F64)
    n_axes = min(n_axes, d)
    proj = empty((n,), F64)
    w = empty((d,), F64)
    idx = torch.arange(d, dtype=F64, device=x.device)
    axes = []
    for a in range(n_axes):
        v = torch.cos((idx + 1.0) * (a + 1.0) * 0.7390851)          # deterministic, generic start
        for it in range(iters + 1):
            for u in axes:
                v = v - torch.dot(v, u) * u
            v = (v / torch.linalg.vector_norm(v)).contiguous()
            if it == iters:
                break
            call("pb200_project_rows", x, is64, n, d, mean, v, proj)
            call("pb200_weighted_col_sums", x, is64, n, d, mean, proj, w)
            v = w.clone()
        axes.append(v)
    V = (torch.stack(axes) * (1.0 - 1e-12)).contiguous()             # |v| <= 1: projections never exceed distances
    _dist.broadcast_(V)                                              # every rank must order identically
    return V


def knn_prepare(x: torch.Tensor, prune: bool = True) -> KnnPrep:
    """Mean-centre, order the points along the Morton curve of their projections on the three
    leading principal axes (work order), box every 128-point tile and build the candidate kernel's
    operands."""
    n, d = x.shape
    if x.dtype not in (F32, F64):
        raise TypeError("kNN input must be float32 or float64")
    d_pad = max(16, round_up(d, 16))
    ld = round_up(n, 256)
    is64 = int(x.dtype == F64)
    mean = empty((d,), F64)
    call("pb200_col_means", x, is64, n, d, mean)
    perm = box_lo = box_hi = None
    if prune and n > 512 and d >= 3:
        V = principal_axes(x, mean)
        proj = empty((3, n), F64)
        for a in range(3):
            call("pb200_project_rows", x, is64, n, d, mean, V[a].contiguous(), proj[a])
        lo = proj.amin(dim=1).cpu()
        hi = proj.amax(dim=1).cpu()
        ext = float((hi - lo).max().item())
        inv_cell = (2 ** 21 - 1) / ext if ext > 0.0 else 1.0
        lib = _native.load()
        nbytes = int(lib.pb200_argsort_scratch_bytes(n))
        scratch = empty((nbytes,), torch.uint8)
        perm, inv = empty((n,), I32), empty((n,), I32)
        call("pb200_morton_order", proj, n, float(lo[0]), float(lo[1]), float(lo[2]), float(inv_cell), perm, inv,
             scratch, nbytes)
        ntiles = ld // 128
        box_lo, box_hi = empty((ntiles, 3), F64), empty((ntiles, 3), F64)
        call("pb200_tile_boxes", proj, perm, n, ntiles, box_lo, box_hi)
        del scratch, inv
    xt = empty((d_pad * ld,), F32)
    yn = empty((ld,), F32)
    xcn = empty((n,), F64)
    xn64 = empty((n,), F64)
    rmax2 = empty((1,), F64)
    call("pb200_knn_prepare", x, is64, n, d, d_pad, ld, mean, perm, xt, yn, xcn, xn64, rmax2)
    xtc = None
    if KNN_ENGINE == "tc" and d <= 128:
        nfl = int(_native.load().pb200_knn_tc_operand_floats(ld, d))
        xtc = empty((nfl,), F32)
        call("pb200_knn_prepare_tc", x, is64, n, d, ld, mean, perm, xtc)
    return KnnPrep(x, perm, xt, yn, xcn, xn64, rmax2, box_lo, box_hi, xtc, mean, n, d, d_pad, ld)


def knn_expanded(prep: KnnPrep, k: int, q0: int = 0, nq: Optional[int] = None,
                 slack: int = 12) -> Tuple[torch.Tensor, torch.Tensor, dict]:
    """k nearest references (self included) of WORK-order rows q0..q0+nq-1, sklearn brute semantics.
    Returns (idx int32 [nq,k] of input-order indices, dist f64 [nq,k], stats); rows are in work order
    (see knn_unpermute).  q0 % 256 == 0."""
    n, d = prep.n, prep.d
    nq = n - q0 if nq is None else nq
    if k > n:
        raise ValueError("k > number of points")
    out_idx = empty((nq, k), I32)
    out_dist = empty((nq, k), F64)
    stats = {"flagged": 0, "k_keep": 0}
    if nq == 0:
        return out_idx, out_dist, stats
    is64 = int(prep.x.dtype == F64)
    if k > KNN_MAX_KEEP:
        # wider than the candidate buffers: exact FP64 brute force for every row
        rows = torch.arange(nq, dtype=I32, device=out_idx.device)
        _knn_exact_rows(prep, rows, nq, q0, k, out_idx, out_dist)
        stats["flagged"] = nq
        return out_idx, out_dist, stats
    k_keep = min(KNN_MAX_KEEP, k + slack)
    stats["k_keep"] = k_keep
    cand = empty((nq, k_keep), I32)
    score = empty((nq, k_keep), F32)
    thr = empty((nq,), F32)
    use_tc = prep.xtc is not None
    cap = int(_native.load().pb200_knn_tc_list_cap()) if use_tc else 64
    lists = empty((round_up(nq, 256) * cap,), I64)
    visited = zeros((1,), I64)
    with stage("knn_candidates"):
        if use_tc:
            call("pb200_knn_candidates_tc", prep.xtc, prep.ld, prep.yn, prep.xcn, prep.box_lo, prep.box_hi, q0, nq,
                 prep.d, k_keep, lists, cand, score, thr, visited)
        else:
            call("pb200_knn_candidates_f32", prep.xt, prep.ld, prep.yn, prep.xcn, prep.box_lo, prep.box_hi, q0, nq,
                 prep.d, prep.d_pad, k_keep, lists, cand, score, thr, visited)
    del lists
    eps_coef = 2.0 * (16 + 4 * 3 * ((d + 7) // 8)) if use_tc else 2.0 * (d + 8)
    stats["engine"] = "tc" if use_tc else "simt"
    flag = empty((nq,), U8)
    n_flagged = empty((1,), I32)
    with stage("knn_rerank"):
        call("pb200_knn_rerank_f64", prep.x, is64, prep.perm, prep.xn64, prep.xcn, prep.rmax2, n, d, q0, nq, cand,
             thr, k_keep, k, eps_coef, out_idx, out_dist, flag, n_flagged)
    nf = int(n_flagged.item())
    stats["flagged"] = nf
    cta_rows = 128 if use_tc else 256
    stats["tiles_visited_frac"] = float(visited.item()) / (max(1, (nq + cta_rows - 1) // cta_rows) * (prep.ld // 128))
    if nf > 0:
        rows = empty((nq,), I32)
        counter = empty((1,), I32)
        call("pb200_knn_flagged_rows", flag, nq, rows, counter)
        with stage("knn_exact_rows"):
            _knn_exact_rows(prep, rows, nf, q0, k, out_idx, out_dist)
    return out_idx, out_dist, stats


def _knn_exact_rows(prep: KnnPrep, rows, n_rows, q0, k, out_idx, out_dist):
    n_ctas = max(1, min(n_rows, 148))
    scratch = empty((n_ctas, prep.n), F64)
    call("pb200_knn_exact_rows_f64", prep.x, int(prep.x.dtype == F64), prep.perm, prep.xn64, prep.n, prep.d, rows,
         n_rows, q0, k, scratch, n_ctas, out_idx, out_dist)


def knn_unpermute(prep: KnnPrep, idx: torch.Tensor, dist: torch.Tensor):
    """Full [n, k] tables with rows in work order -> rows in input order."""
    if prep.perm is None:
        return idx, dist
    oi, od = torch.empty_like(idx), torch.empty_like(dist)
    call("pb200_knn_unpermute", idx, dist, idx.shape[0], idx.shape[1], prep.perm, oi, od)
    return oi, od


def knn_direct(data: torch.Tensor, k: int, q0: int = 0, nq: Optional[int] = None,
               query: Optional[torch.Tensor] = None, qrows: Optional[torch.Tensor] = None):
    """Exact FP64 direct-difference kNN (kd-tree equivalent), m <= 16 columns."""
    n, m = data.shape
    if data.dtype != F64:
        raise TypeError("knn_direct needs float64 data")
    query = data if query is None else query
    if qrows is not None:
        nq = qrows.numel()
    elif nq is None:
        nq = query.shape[0] - q0
    out_idx = empty((nq, k), I32)
    out_dist = empty((nq, k), F64)
    with stage("knn_direct"):
        call("pb200_knn_direct_f64", data, n, m, query, qrows, q0, nq, k, out_idx, out_dist)
    return out_idx, out_dist


KNN_BALANCE_MIN_ROWS = 262144
# sampled query tiles fetch this many reference tiles, then only count the tiles their pruning bound still admits
# (0: they run to the end -- one CTA lifetime, ~2.7 ms at 1 M cells, against ~0.1 ms)
KNN_BALANCE_VISIT_LIMIT = int(os.environ.get("PB200_KNN_BALANCE_LIMIT", "24"))


def cuts_at_equal_work(work: np.ndarray, parts: int, n: int) -> List[int]:
    """parts + 1 row boundaries over range(n) from the work of every 128-row tile: interior boundaries are whole pairs
    of tiles (multiples of 256 rows, the alignment knn_expanded needs), monotone, and split the cumulative work into
    equal shares as closely as that granularity allows."""
    n_t = len(work)
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(work, dtype=np.float64))])
    cuts = [0]
    for r in range(1, parts):
        t = int(np.searchsorted(cum, cum[-1] * r / parts))
        t = max(cuts[-1] // 128, min(n_t, (t + 1) // 2 * 2))
        cuts.append(min(n, t * 128))
    cuts.append(n)
    return cuts


def balanced_knn_bounds(prep: KnnPrep, k: int, slack: int = 12):
    """Row boundaries (multiples of 256, in work order) that give every rank the same share of the candidate
    kernel's work, or None (one rank, small input, SIMT engine).  The work of a 128-row query tile is the number of
    reference tiles it visits -- 5 % to 25 % of them depending on where along the trajectory it lies -- so equal
    row counts leave the slowest rank ~20 % behind the mean (measured: 85.7 ms against 70.5 at 2 GPUs, 26.2 against
    17.6 at 8).  One wave of sampled query tiles (every ceil(T/148)-th) is run through the kernel itself
    (pb200_knn_tc_visit_counts) for KNN_BALANCE_VISIT_LIMIT tiles each and then counts the tiles its pruning bound
    still admits; the counts are interpolated along the work order and the rows are cut at equal cumulative counts.  Rank 0's cut is broadcast: the counts depend on the timing of the pruning bound."""
    from . import _dist

    rank, size = _dist.world()
    n = prep.n
    if size == 1 or prep.xtc is None or n < KNN_BALANCE_MIN_ROWS or k > KNN_MAX_KEEP:
        return None
    k_keep = min(KNN_MAX_KEEP, k + slack)
    n_qt = (n + 127) // 128
    stride = max(1, -(-n_qt // 148))
    sample = torch.arange(stride // 2, n_qt, stride, dtype=I32, device=prep.yn.device)
    ns = int(sample.numel())
    counts = zeros((ns,), I32)
    cand = empty((ns * 128, k_keep), I32)
    score = empty((ns * 128, k_keep), F32)
    thr = empty((ns * 128,), F32)
    with stage("knn_balance"):
        call("pb200_knn_tc_visit_counts", prep.xtc, prep.ld, prep.yn, prep.xcn, prep.box_lo, prep.box_hi, n, prep.d,
             k_keep, sample, ns, KNN_BALANCE_VISIT_LIMIT, cand, score, thr, counts)
        work = np.interp(np.arange(n_qt), sample.cpu().numpy().astype(np.float64),
                         np.maximum(counts.cpu().numpy().astype(np.float64), 1.0))
    cuts = cuts_at_equal_work(work, size, n)
    b = torch.tensor(cuts, dtype=torch.int64, device=prep.yn.device)
    _dist.broadcast_(b)
    return [int(v) for v in b.tolist()]


def knn_like_sklearn(x: torch.Tensor, n_neighbors: int, shard: bool = False):
    """Full self-including neighbour table, dispatching exactly as sklearn's algorithm='auto' would
    (see sklearn_uses_brute).  With shard=True the query rows are split across the ranks of the
    process group (contiguous blocks of the work order) and the row shards all-gathered."""
    from . import _dist

    n, d = x.shape
    q0, nq = _dist.row_shard(n, 256) if shard else (0, n)
    if sklearn_uses_brute(n, d, n_neighbors):
        prep = knn_prepare(x)
        bounds = balanced_knn_bounds(prep, n_neighbors) if shard else None
        if bounds is not None:
            rank = _dist.world()[0]
            q0, nq = bounds[rank], bounds[rank + 1] - bounds[rank]
        idx, dist, stats = knn_expanded(prep, n_neighbors, q0, nq)
        if shard:
            idx, dist = _dist.allgather_rows(idx, dist, n, 256, bounds)
            stats["shard_rows"] = nq
        idx, dist = knn_unpermute(prep, idx, dist)
        stats["method"] = "expanded"
        if prep.perm is not None:
            stats["perm"] = prep.perm          # work order (callers pop it: locality order for the eigensolver's SpMV)
        return idx, dist, stats
    data = x if x.dtype == F64 else x.to(F64)
    idx, dist = knn_direct(data.contiguous(), n_neighbors, q0, nq)
    if shard:
        idx, dist = _dist.allgather_rows(idx, dist, n, 256)
    return idx, dist, {"method": "direct", "flagged": 0}


# ---------------------------------------------------------------------------
# CSR construction
# ---------------------------------------------------------------------------
@dataclass
class DeviceCSR:
    indptr: torch.Tensor    # int32 [n+1]
    indices: torch.Tensor   # int32 [nnz]
    data: torch.Tensor      # f64 [nnz]
    n: int
    order: Optional[torch.Tensor] = None    # int32 [n]: a locality-preserving order of the rows (order[i] = row at position i)
    symmetric: bool = False                 # built here as W + W^T (and scaled symmetrically): no need to check

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())


def symmetrize(idx: torch.Tensor, w: torch.Tensor, col0: int, mode: int, drop_self: bool) -> DeviceCSR:
    """Union-pattern symmetrisation of the arc list (idx[n,kc], w[n,kc-col0]):
    mode 0 -> w_ij + w_ji, mode 1 -> min(w_ij, w_ji).  Sorted columns."""
    n, kc = idx.shape
    keff = kc - col0
    indeg = empty((n,), I32)
    ub = empty((n,), I32)
    cursor = empty((n,), I32)
    off = empty((n + 1,), I64)
    bsum = empty((n // 1024 + 3,), I64)
    tcol = empty((2 * n * keff,), I32)
    tval = empty((2 * n * keff,), F64)
    row_nnz = empty((n,), I32)
    indptr = empty((n + 1,), I32)
    call("pb200_symmetrize_phase1", idx, w, n, kc, col0, mode, int(drop_self), indeg, ub, cursor, off, bsum,
         tcol, tval, row_nnz, indptr)
    nnz = int(indptr[n].item())
    indices = empty((nnz,), I32)
    data = empty((nnz,), F64)
    call("pb200_symmetrize_phase2", n, off, indptr, tcol, tval, indices, data)
    return DeviceCSR(indptr, indices, data, n)


def adaptive_kernel(idx: torch.Tensor, dist: torch.Tensor, knn: int, alpha: float) -> DeviceCSR:
    """utils.py:408-417,440-446 on the device from a neighbour table that includes self."""
    n, kc = idx.shape
    col0 = 0
    if kc > 1:
        flag = empty((1,), I32)
        call("pb200_knn_self_first", idx, n, kc, flag)
        col0 = 1 if int(flag.item()) == 0 else 0
    keff = kc - col0
    ka = max(1, min(int(math.floor(knn / 3)), keff))
    w = empty((n, keff), F64)
    with stage("kernel_build"):
        call("pb200_adaptive_weights", dist, n, kc, col0, ka, w)
        kern = symmetrize(idx, w, col0, 0, False)
        if alpha > 0:
            deg = empty((n,), F64)
            call("pb200_csr_row_sums", kern.indptr, kern.data, n, deg)
            call("pb200_vec_transform", deg, n, 1, float(-alpha))
            out = empty((kern.nnz,), F64)
            call("pb200_csr_scale", kern.indptr, kern.indices, kern.data, n, deg, deg, out)
            kern = DeviceCSR(kern.indptr, kern.indices, out, n)
    kern.symmetric = True       # w_ij + w_ji is computed once per pair; D^-a K D^-a scales both mirror entries alike
    return kern


def csr_from_host(m) -> DeviceCSR:
    """scipy CSR -> device (int32 indices, float64 values, sorted columns)."""
    m = m.tocsr()
    if not m.has_sorted_indices:
        m = m.sorted_indices()
    if m.nnz >= 2 ** 31:
        raise ValueError("kernel has >= 2^31 stored entries")
    return DeviceCSR(to_device(m.indptr.astype(np.int32, copy=False)),
                     to_device(m.indices.astype(np.int32, copy=False)),
                     to_device(m.data.astype(np.float64, copy=False)), m.shape[0])


def csr_to_host(c: DeviceCSR, data: Optional[torch.Tensor] = None):
    from scipy.sparse import csr_matrix

    vals, indices, indptr = to_host(c.data if data is None else data, c.indices, c.indptr)
    return csr_matrix((vals, indices, indptr), shape=(c.n, c.n))


# ---------------------------------------------------------------------------
# eigensolver
# ---------------------------------------------------------------------------
class NotSymmetricError(NotImplementedError):
    pass


def transition_and_symmetric(kern: DeviceCSR):
    """Row sums D, T = D^-1 K values, S = D^-1/2 K D^-1/2 values, D^-1/2, D^1/2."""
    n = kern.n
    deg = empty((n,), F64)
    call("pb200_csr_row_sums", kern.indptr, kern.data, n, deg)
    inv = deg.clone()
    call("pb200_vec_transform", inv, n, 0, 0.0)
    tvals = empty((kern.nnz,), F64)
    call("pb200_csr_scale", kern.indptr, kern.indices, kern.data, n, inv, None, tvals)
    dis = deg.clone()
    call("pb200_vec_transform", dis, n, 2, 0.0)
    svals = empty((kern.nnz,), F64)
    call("pb200_csr_scale", kern.indptr, kern.indices, kern.data, n, dis, dis, svals)
    dsq = deg.clone()
    call("pb200_vec_transform", dsq, n, 3, 0.0)
    return tvals, svals, dis, dsq


# Gram-Schmidt pass of the Lanczos recurrence on an FP32 shadow of the basis (see pb200_lanczos_steps)
LANCZOS_FP32_SHADOW = os.environ.get("PB200_LANCZOS_SHADOW", "1") != "0"
# eigensolver on the matrix relabelled into the kNN stage's Morton order (x gathers 9 % faster: 0.305 vs 0.336 ms
# per SpMV at 1M cells); PB200_SPMV_WINDOW=1 additionally switches to the shared-memory-window SpMV kernel, which
# is correct but slower so far (0.43-0.59 ms: 16 warps per SM do not cover the HBM latency of the CSR stream)
LANCZOS_WINDOWED = os.environ.get("PB200_LANCZOS_WINDOWED", "1") != "0"
# Gram-Schmidt against the whole basis only when the omega recurrence asks for it (partial reorthogonalisation,
# pb200_lanczos_steps_pro): ~1 step in 8 instead of every step
LANCZOS_PRO = os.environ.get("PB200_LANCZOS_PRO", "1") != "0"
LANCZOS_INITIAL_BASIS = 320       # vectors allocated up front; the basis is re-allocated (doubling) beyond that
# several ranks: every rank multiplies its slice of the rows, the slices of w are all-gathered, and the (cheap,
# deterministic) vector part of the step is repeated on every rank -- the basis is never exchanged
LANCZOS_SHARDED = os.environ.get("PB200_LANCZOS_SHARDED", "1") != "0"
LANCZOS_SHARD_MIN_ROWS = 65536
# product over the nonzero stream (one CTA per 1024 nonzeros, no separate kernel for the hub rows)
SPMV_STREAM = os.environ.get("PB200_SPMV_STREAM", "1") != "0"
SPMV_WINDOW_KERNEL = os.environ.get("PB200_SPMV_WINDOW", "0") == "1"


def lanczos_topk(kern: DeviceCSR, svals: torch.Tensor, dis: torch.Tensor, start: torch.Tensor, k: int,
                 tol: float = 1e-10, batch: int = 20, max_basis: Optional[int] = None):
    """Largest-magnitude k eigenpairs of S (= those of T) by symmetric Lanczos with
    full re-orthogonalisation.  Returns (eigenvalues desc [k] numpy, U device [n,k] =
    unit-L2 columns of D^-1/2 u, info dict).

    The small tridiagonal eigenproblem (m x m, m ~ a few hundred) is solved on the
    host between batches of device steps -- control logic, like ARPACK's own
    dseupd; all O(N) work stays on the device."""
    from scipy.linalg import eigh_tridiagonal

    n, nnz = kern.n, kern.nnz
    if not (0 < k < n):
        raise ValueError("need 0 < k < n eigenpairs")
    # With a locality-preserving row order at hand (the kNN stage's Morton order) the solver runs on the
    # relabelled matrix P S P^T: same eigenvalues, eigenvectors permuted back at the end; the SpMV's gathers of
    # x[col] then fall into a narrow band around the row.
    perm = kern.order if (LANCZOS_WINDOWED and kern.order is not None and n >= 4 * 24576) else None
    indptr, indices = kern.indptr, kern.indices
    if perm is not None:
        pl = perm.long()
        inv = torch.empty_like(perm)
        inv[pl] = torch.arange(n, dtype=perm.dtype, device=perm.device)
        lens = (kern.indptr[1:] - kern.indptr[:-1])[pl]
        indptr = torch.zeros((n + 1,), dtype=I32, device=perm.device)
        indptr[1:] = torch.cumsum(lens, 0)
        indices = empty((nnz,), I32)
        svals_p = empty((nnz,), F64)
        call("pb200_csr_permute", kern.indptr, kern.indices, svals, n, perm, inv, indptr, indices, svals_p)
        svals = svals_p
        start = start[pl].contiguous()
    if max_basis is None:
        # keep the basis under ~40 GB
        max_basis = int(min(n, max(4 * k + 40, min(2048, (40 << 30) // (8 * n)))))
    max_basis = min(max_basis, n)
    # hub rows of the symmetrised kNN graph: listed once, walked by a whole warp each in every SpMV
    lens = indptr[1:] - indptr[:-1]
    long_rows = torch.nonzero(lens > int(_native.load().pb200_spmv_long_row())).to(I32).reshape(-1).contiguous()
    n_long = int(long_rows.numel())
    if n_long == 0:
        long_rows = None
    from . import _dist
    rank, world_size = _dist.world()
    sharded = bool(LANCZOS_SHARDED and LANCZOS_PRO and world_size > 1 and n >= LANCZOS_SHARD_MIN_ROWS)
    row0, row1 = 0, n
    if sharded:
        import torch.distributed as tdist
        chunk = round_up(-(-n // world_size), 16)
        row0, row1 = min(rank * chunk, n), min((rank + 1) * chunk, n)
    stream_spmv = bool(SPMV_STREAM and LANCZOS_PRO and n >= 32768)
    if stream_spmv:
        tile = int(_native.load().pb200_csr_spmv_stream_tile())
        n_tiles = -(-nnz // tile)
        tile_lo = empty((n_tiles + 1,), I32)
        st_head, st_tail = empty((n_tiles,), F64), empty((n_tiles,), F64)
        call("pb200_csr_spmv_stream_prepare", indptr, n, nnz, tile_lo)
        if sharded:
            e0, e1 = (int(v) for v in indptr[torch.tensor([row0, row1], device=indptr.device)].tolist())
            tile0, tile1 = e0 // tile, -(-e1 // tile)
        else:
            tile0, tile1 = 0, n_tiles
    elif sharded:
        if long_rows is not None:
            mine = long_rows[(long_rows >= row0) & (long_rows < row1)].contiguous()
            long_mine, n_long_mine = (mine, int(mine.numel())) if mine.numel() else (None, 0)
        else:
            long_mine, n_long_mine = None, 0
    ld = n
    nblk = n // 2048 + 1
    # the basis grows on demand (the cap is ~2000 vectors = 16 GB at 1 M cells; 200 are used there)
    cap = min(max_basis, max(LANCZOS_INITIAL_BASIS, batch))
    V = empty((cap + 1, ld), F64)
    V32 = empty((cap + 1, ld), F32) if (LANCZOS_FP32_SHADOW and not LANCZOS_PRO) else None
    alpha = zeros((max_basis + 1,), F64)
    beta = zeros((max_basis + 2,), F64)
    w = empty((world_size * chunk,), F64) if sharded else empty((n,), F64)
    h = empty((max_basis + 1,), F64)
    partial = empty(((max_basis + 1) * nblk,), F64)
    tmp = empty((1,), F64)
    call("pb200_lanczos_init", start, n, V, V32, partial, tmp)
    omstride = max_basis + 2
    om = zeros((3 * omstride,), F64)
    om[0] = 1.0
    gate = zeros((2,), I32)
    counters = zeros((1,), I32)
    eps_eff = 2.2e-16 * math.sqrt(n)
    m = 0
    info = {"converged": False}
    theta = yvec = None
    step_batch = batch
    lagged = bool(stream_spmv and batch > 5)          # large problems: convergence checks one batch behind the device
    pending = None
    while m < max_basis:
        steps = min(step_batch, max_basis - m)
        if m + steps > cap:
            cap = min(max_basis, max(2 * cap, m + steps))
            V_new = empty((cap + 1, ld), F64)
            V_new[:m + 1].copy_(V[:m + 1])
            V = V_new
            if V32 is not None:
                V32_new = empty((cap + 1, ld), F32)
                V32_new[:m + 1].copy_(V32[:m + 1])
                V32 = V32_new
            del V_new
        with stage("lanczos_steps"):
            if sharded or stream_spmv:
                w_mine = w[rank * chunk:(rank + 1) * chunk] if sharded else None
                for j in range(m, m + steps):
                    if stream_spmv:
                        call("pb200_csr_spmv_stream_f64", indptr, indices, svals, n, nnz, tile_lo, st_head, st_tail,
                             V[j], w, row0, row1, tile0, tile1)
                    else:
                        call("pb200_csr_spmv_rows_f64", indptr, indices, svals, n, nnz, row0, row1, long_mine,
                             n_long_mine, V[j], w)
                    if sharded:
                        tdist.all_gather_into_tensor(w, w_mine)
                    call("pb200_lanczos_post_pro", n, V, ld, j, alpha, beta, w, h, partial, tmp, om, omstride, gate,
                         counters, eps_eff, 1.5e-8)
            elif LANCZOS_PRO:
                call("pb200_lanczos_steps_pro", indptr, indices, svals, n, nnz, long_rows, n_long, V, ld, V32,
                     int(perm is not None and SPMV_WINDOW_KERNEL), m, steps,
                     alpha, beta, w, h, partial, tmp, om, omstride, gate, counters, eps_eff, 1.5e-8)
            else:
                call("pb200_lanczos_steps", indptr, indices, svals, n, nnz, long_rows, n_long, V, ld, V32,
                     int(perm is not None and SPMV_WINDOW_KERNEL), m, steps,
                     alpha, beta, w, h, partial, tmp)
        m += steps
        if m < k + 1 and m < max_basis:
            continue
        looks = []
        if lagged and step_batch == batch and m < max_basis:
            # Far from convergence (20-step batches): the coefficients of this batch go to the host asynchronously and
            # the host looks at the PREVIOUS batch while the device runs this one -- the m x m eigenproblem (~1 ms at
            # m = 200) and the two read-backs no longer stall the stream a dozen times per solve.
            a_h = torch.empty((m,), dtype=F64, pin_memory=True)
            b_h = torch.empty((m,), dtype=F64, pin_memory=True)
            a_h.copy_(alpha[:m], non_blocking=True)
            b_h.copy_(beta[1:m + 1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            prev, pending = pending, (m, a_h, b_h, ev)
            if prev is None:
                continue
            looks.append(prev)
        looks.append(None)                            # the batch just run (skipped while the previous one looks far off)
        done = False
        for look in looks:
            if look is not None:
                m_chk, a_t, b_t, ev_prev = look
                ev_prev.synchronize()
                a, b = a_t.numpy(), b_t.numpy()
            else:
                if len(looks) == 2 and step_batch == batch:
                    break                             # still far from convergence: keep running one batch ahead
                pending = None
                m_chk = m
                a = alpha[:m].cpu().numpy()
                b = beta[1:m + 1].cpu().numpy()          # b[j] = beta_{j+1}
            theta_all, y_all = eigh_tridiagonal(a, b[:m_chk - 1])
            order = np.argsort(-np.abs(theta_all))[:k]
            resid = np.abs(b[m_chk - 1] * y_all[m_chk - 1, order])
            theta, yvec = theta_all[order], y_all[:, order]
            info.update(steps=m_chk, max_resid=float(resid.max()))
            info.setdefault("history", []).append((m_chk, float(resid.max())))
            # the residuals fall by orders of magnitude over the last ten steps (1e-4 -> 1e-9 -> 1e-15 at 1 M cells):
            # once they are small, look every five steps instead of every twenty (and at the batch just run)
            if resid.max() < 1e-3 * max(1e-3, np.abs(theta).max()):
                step_batch = max(1, min(batch, 5))
            exhausted = (m_chk >= n) or (b[m_chk - 1] <= 1e-14 * max(1.0, np.abs(theta_all).max()))
            done = bool(np.all(resid <= tol * np.maximum(np.abs(theta), 1e-3)) or exhausted)
            if sharded:
                # the ranks run identical kernels on identical data, so they agree; rank 0's word is taken anyway,
                # because a disagreement would leave the all-gathers of the next batch unmatched
                ctl = torch.tensor([int(done), step_batch], dtype=torch.int64, device=V.device)
                _dist.broadcast_(ctl)
                done, step_batch = bool(ctl[0].item()), int(ctl[1].item())
            if done:
                break
        if done:
            m = m_chk                                # steps run beyond the batch that was looked at are not used
        if done:
            info["converged"] = True
            break
    info["reorthogonalised_steps"] = int(counters.item()) if LANCZOS_PRO else m
    info["row_sharded_over"] = world_size if sharded else 1
    info["spmv"] = "nonzero stream" if stream_spmv else "row kernels"
    if not info["converged"]:
        raise RuntimeError(
            f"Lanczos did not converge in {m} steps (max residual {info.get('max_resid')}); "
            "raise max_basis")
    srt = np.argsort(-theta)
    theta, yvec = theta[srt], np.ascontiguousarray(yvec[:, srt])
    U = empty((n, k), F64)
    with stage("ritz_vectors"):
        if perm is None:
            call("pb200_ritz_vectors", V, ld, m, to_device(yvec), k, dis, U, n)
        else:
            Up = empty((n, k), F64)
            call("pb200_ritz_vectors", V, ld, m, to_device(yvec), k, dis[perm.long()].contiguous(), Up, n)
            call("pb200_unpermute_rows_f64", Up, n, k, perm, U)
        part2 = empty((k * (n // 4096 + 1),), F64)
        sumsq = empty((k,), F64)
        call("pb200_normalize_columns", U, n, k, part2, sumsq)
    return theta, U, info


def arnoldi_topk(kern: DeviceCSR, tvals: torch.Tensor, start: torch.Tensor, k: int, tol: float = 1e-10,
                 batch: int = 20, max_basis: Optional[int] = None):
    """Largest-magnitude k eigenpairs of the row-stochastic T = D^-1 K of a kernel that is NOT symmetric (the
    reference hands any kernel to ARPACK's general solver, utils.py:484) by Arnoldi with two Gram-Schmidt passes,
    un-restarted.  Returns (real parts of the eigenvalues, descending [k]; U [n, k] = real parts of the Ritz
    vectors, unit L2 columns; info) -- the reference keeps the real parts too (utils.py:485-486)."""
    n, nnz = kern.n, kern.nnz
    if not (0 < k < n):
        raise ValueError("need 0 < k < n eigenpairs")
    if max_basis is None:
        max_basis = int(min(n - 1, max(4 * k + 40, min(1024, (20 << 30) // (8 * n)))))
    ld = n
    nblk = n // 2048 + 1
    V = empty((max_basis + 1, ld), F64)
    ldh = max_basis + 1
    H = zeros((max_basis, ldh), F64)                         # column-major: H[j] is column j
    w, h, h2 = empty((n,), F64), empty((max_basis + 1,), F64), empty((max_basis + 1,), F64)
    partial = empty(((max_basis + 1) * nblk,), F64)
    tmp = empty((1,), F64)
    call("pb200_lanczos_init", start, n, V, None, partial, tmp)
    m, info = 0, {"converged": False, "solver": "arnoldi"}
    lam = yv = None
    while m < max_basis:
        steps = min(batch, max_basis - m)
        with stage("lanczos_steps"):
            call("pb200_arnoldi_steps", kern.indptr, kern.indices, tvals, n, nnz, V, ld, m, steps, H, ldh, w, h, h2,
                 partial, tmp)
        m += steps
        if m < k + 1 and m < max_basis:
            continue
        Hh = H[:m].cpu().numpy().T                           # [ldh, m]: rows 0..m
        vals, vecs = np.linalg.eig(Hh[:m, :m])
        order = np.argsort(-np.abs(vals))[:k]
        resid = np.abs(Hh[m, m - 1]) * np.abs(vecs[m - 1, order])
        lam, yv = vals[order], vecs[:, order]
        info.update(steps=m, max_resid=float(resid.max()))
        if np.all(resid <= tol * np.maximum(np.abs(lam), 1e-3)) or Hh[m, m - 1] <= 1e-14:
            info["converged"] = True
            break
    if not info["converged"]:
        raise RuntimeError(f"Arnoldi did not converge in {m} steps (max residual {info.get('max_resid')})")
    lam_r = np.real(lam)
    srt = np.argsort(-lam_r)
    lam_r, yr = lam_r[srt], np.ascontiguousarray(np.real(yv[:, srt]))
    U = empty((n, k), F64)
    call("pb200_ritz_vectors", V, ld, m, to_device(yr), k, None, U, n)
    part2 = empty((k * (n // 4096 + 1),), F64)
    sumsq = empty((k,), F64)
    call("pb200_normalize_columns", U, n, k, part2, sumsq)
    return lam_r, U, info


def check_symmetric(kern: DeviceCSR, tol: float = 1e-12) -> bool:
    flag = empty((1,), I32)
    call("pb200_csr_is_symmetric", kern.indptr, kern.indices, kern.data, kern.n, tol, flag)
    return int(flag.item()) == 0
