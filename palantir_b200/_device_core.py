"""Device-side orchestration of the run_palantir stages (scaling, waypoint sampling,
kNN graph, reachability, multi-source shortest paths, refinement, Markov chain,
absorption solve, projection).  CUDA tensors in, CUDA tensors out; see core.py for
the host mirror of the reference API.
"""
from __future__ import annotations

import math
import warnings
from typing import Optional

import numpy as np
import torch

from . import _device as dv
from . import _dist, _native
from ._device import F64, I32, I64, call, empty, stage, zeros

U32 = torch.int32  # bitmaps / counters are passed as raw 32-bit words


# ---------------------------------------------------------------------------
# scaling, boundaries, sampling
# ---------------------------------------------------------------------------
def col_minmax(a: torch.Tensor):
    """(min, argmin, max, argmax) per column, first occurrence on ties."""
    n, m = a.shape
    nblk = min(1024, (n + 255) // 256)
    omin, omax = empty((m,), F64), empty((m,), F64)
    oimin, oimax = empty((m,), I64), empty((m,), I64)
    pval = empty((2 * m * nblk,), F64)
    pidx = empty((2 * m * nblk,), I64)
    call("pb200_col_minmax", a, n, m, omin, oimin, omax, oimax, pval, pidx)
    return omin, oimin, omax, oimax


def minmax_scale(ms: torch.Tensor) -> torch.Tensor:
    n, m = ms.shape
    mn, _, mx, _ = col_minmax(ms)
    out = empty((n, m), F64)
    call("pb200_minmax_apply", ms, n, m, mn, mx, out)
    return out


def maxmin_sampling(data: torch.Tensor, first: np.ndarray, iters: int) -> np.ndarray:
    """[m, iters] waypoint positions: column c starts at first[c] (drawn on the host with
    the reference's RNG calls) and adds iters-1 farthest points."""
    n, m = data.shape
    nblk = min(592, (n + 255) // 256)
    out = empty((m, iters), I64)
    vt = empty((m * n,), F64)
    md = empty((m * n,), F64)
    cur = empty((m,), I64)
    pv = empty((m * nblk,), F64)
    pi = empty((m * nblk,), I64)
    done = empty((m,), I32)
    with stage("maxmin_sampling"):
        call("pb200_maxmin_sampling", data, n, m, dv.to_device(first.astype(np.int64)), iters, out, vt, md, cur,
             pv, pi, done)
    return out.cpu().numpy()


def gather_rows(a: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    m = 1 if a.dim() == 1 else a.shape[1]
    out = empty((rows.numel(),) if a.dim() == 1 else (rows.numel(), m), F64)
    call("pb200_gather_rows_f64", a, m, rows, rows.numel(), out)
    return out


# ---------------------------------------------------------------------------
# kNN graph, reachability, shortest paths
# ---------------------------------------------------------------------------
def knn_table(data: torch.Tensor, knn: int):
    """Self-including neighbour table of all cells in the (low-dimensional) multiscale
    space, rows sharded across ranks."""
    n = data.shape[0]
    if knn > n:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {knn}, n_samples_fit = {n}")
    return dv.knn_like_sklearn(data, knn, shard=True)


def locality_order(dev: torch.Tensor, want_sorted: bool = True):
    """Sort the cells along the first column.  Returns (perm, inv, sorted copy):
    sorted[i] = dev[perm[i]], inv[perm[i]] = i."""
    n, m = dev.shape
    lib = _native.load()
    nbytes = int(lib.pb200_argsort_scratch_bytes(n))
    scratch = empty((nbytes,), torch.uint8)
    perm, inv = empty((n,), I32), empty((n,), I32)
    call("pb200_argsort_col_f64", dev, n, m, 0, perm, inv, scratch, nbytes)
    if not want_sorted:
        return perm, inv, None
    out = empty((n, m), F64)
    call("pb200_permute_rows_f64", dev, n, m, perm, out)
    return perm, inv, out


def morton_order(dev: torch.Tensor):
    """Order the cells along the Morton curve of their first three columns (the leading multiscale
    components: the trajectory is a thin tube in that space, so consecutive cells are neighbours and
    the 128-cell tiles have small bounding boxes).  Returns (perm, inv, ordered copy):
    ordered[i] = dev[perm[i]], inv[perm[i]] = i."""
    n, m = dev.shape
    lib = _native.load()
    mm = min(3, m)
    proj = zeros((3, n), F64)
    proj[:mm] = dev[:, :mm].t()
    lo = proj.amin(dim=1).cpu()
    hi = proj.amax(dim=1).cpu()
    ext = float((hi - lo).max().item())
    inv_cell = (2 ** 21 - 1) / ext if ext > 0.0 else 1.0
    nbytes = int(lib.pb200_argsort_scratch_bytes(n))
    scratch = empty((nbytes,), torch.uint8)
    perm, inv = empty((n,), I32), empty((n,), I32)
    call("pb200_morton_order", proj, n, float(lo[0]), float(lo[1]), float(lo[2]), float(inv_cell), perm, inv,
         scratch, nbytes)
    out = empty((n, m), F64)
    call("pb200_permute_rows_f64", dev, n, m, perm, out)
    return perm, inv, out


def knn_table_boxed(work: torch.Tensor, knn: int):
    """Self-including neighbour table of cells stored in a locality-preserving order (exact, pruned by
    the bounding boxes of the 128-cell tiles); indices are stored positions; rows sharded across ranks."""
    n, m = work.shape
    if knn > n:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {knn}, n_samples_fit = {n}")
    q0, nq = _dist.row_shard(n, 256)
    ntiles = (n + 127) // 128
    box_lo, box_hi = empty((ntiles, m), F64), empty((ntiles, m), F64)
    idx = empty((nq, knn), I32)
    dist = empty((nq, knn), F64)
    scanned = zeros((1,), I64)
    with stage("knn_direct"):
        call("pb200_tile_boxes_md", work, n, m, box_lo, box_hi)
        call("pb200_knn_boxed_f64", work, n, m, q0, nq, knn, box_lo, box_hi, idx, dist, scanned)
    idx, dist = _dist.allgather_rows(idx, dist, n, 256)
    return idx, dist, {"method": "direct-boxed", "flagged": 0, "scanned": scanned}


def knn_table_sorted(dsort: torch.Tensor, knn: int):
    """Self-including neighbour table of cells already sorted along column 0 (exact, pruned);
    indices are sorted positions; rows sharded across ranks."""
    n, m = dsort.shape
    if knn > n:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {knn}, n_samples_fit = {n}")
    q0, nq = _dist.row_shard(n, 256)
    idx = empty((nq, knn), I32)
    dist = empty((nq, knn), F64)
    scanned = zeros((1,), I64)
    with stage("knn_direct"):
        call("pb200_knn_sorted_f64", dsort, n, m, q0, nq, knn, idx, dist, scanned)
    idx, dist = _dist.allgather_rows(idx, dist, n, 256)
    return idx, dist, {"method": "direct-sorted", "flagged": 0, "scanned": scanned}


def unpermute_rows(a: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """out[perm[i]] = a[i] (rows)."""
    n = a.shape[0]
    m = 1 if a.dim() == 1 else a.shape[1]
    out = torch.empty_like(a)
    if n == 0 or m == 0:                       # e.g. fate probabilities when no terminal state was found
        return out
    call("pb200_unpermute_rows_f64", a.contiguous(), n, m, perm, out)
    return out


def undirected_graph(idx: torch.Tensor, dist: torch.Tensor) -> dv.DeviceCSR:
    """Symmetric CSR with min(w_ij, w_ji) on the union pattern, self loops dropped: the graph
    csgraph.dijkstra(directed=False) effectively walks (core.py:362)."""
    with stage("graph_build"):
        return dv.symmetrize(idx, dist.contiguous(), 0, 1, True)


def unreachable_count(g: dv.DeviceCSR, start: int):
    label = empty((g.n,), I32)
    count = empty((1,), I32)
    call("pb200_graph_components", g.indptr, g.indices, g.n, label)
    call("pb200_count_other_label", label, g.n, start, count)
    return int(count.item()), label


# "arcrec": arc records + L2 prefetch with 256/512 threads per source when few sources are resident
# (measured at 1 M cells: 197 vs 262 ms at 150 sources, 245 vs 271 at 300); from ~450 sources on the CSR kernel with
# 128 threads per source wins (281 vs 291 ms at 599 sources, 363 vs 462 at 1198) and needs half the scratch.
# "csr": always CSR.
SSSP_MODE = "arcrec"
SSSP_THREADS = None        # None: chosen from the number of sources


def _sssp_threads(n_sources: int) -> int:
    """Threads per source: as many as possible while every source stays resident (one wave)."""
    if SSSP_THREADS is not None:
        return SSSP_THREADS
    for t in (512, 256):
        if n_sources <= (1280 // t) * 148:
            return t
    return 128


def _default_delta(g: dv.DeviceCSR) -> float:
    # bucket width: measured 2x .. 5x the mean arc weight at 1198 sources / 1 M cells: 393, 390 (2.5x), 397, 403,
    # 410 ms -- fewer re-relaxations against more threshold advances, a flat optimum
    delta = 2.5 * float(g.data.mean().item()) if g.nnz else 1.0
    return delta if delta > 0.0 else 1.0


def sssp_per_source(g: dv.DeviceCSR, sources: torch.Tensor, delta: Optional[float] = None,
                    want_stats: bool = False):
    """D[S, n] shortest-path distances from each source (unreachable = +inf): one CTA per source walking the
    whole graph (sssp.cu).  Bit-equal to a heap Dijkstra."""
    n, S = g.n, int(sources.numel())
    if delta is None:
        delta = _default_delta(g)
    D = empty((S, n), F64)
    lib = _native.load()
    stats = zeros((S, 3), I64) if want_stats else None
    nwords = (n + 31) // 32
    def exact_rerun():
        # both default kernels track list membership in shared-memory tag tables; a duplicate that slips through a
        # collision is harmless, but a list may then in theory outgrow its N entries: checked after every call, never
        # observed; the global-bitmap form of the kernel has no such case
        nc = lib.pb200_sssp_num_ctas(S)
        q2 = empty((nc * 3 * n,), I32)
        b2 = empty((nc * 3 * nwords,), U32)
        with stage("sssp"):
            call("pb200_sssp_multi", g.indptr, g.indices, g.data, n, sources, S, float(delta), 1, D, q2, b2, stats)

    if SSSP_MODE == "arcrec" and (SSSP_THREADS is not None or S <= 444):
        threads = _sssp_threads(S)
        n_ctas = lib.pb200_sssp_arcrec_ctas(S, threads)
        arcs = empty(((g.nnz + 64) * 2,), F64)                     # 16-byte records
        with stage("sssp"):
            call("pb200_sssp_build_arcs", g.indptr, g.indices, g.data, g.nnz, arcs)
            q = empty((n_ctas * 3 * n * 2,), I32)
            bits = empty((n_ctas * 3 * nwords,), U32)
            call("pb200_sssp_multi_arcrec", g.indptr, arcs, n, sources, S, float(delta), threads, D, q, bits, stats)
        del q, bits, arcs
        if int(lib.pb200_sssp_overflowed(_native.stream_ptr())):
            exact_rerun()
        return D, stats
    n_ctas = lib.pb200_sssp_num_ctas(S)
    q = empty((n_ctas * 3 * n,), I32)
    bits = empty((n_ctas * 3 * nwords,), U32)
    with stage("sssp"):
        call("pb200_sssp_multi", g.indptr, g.indices, g.data, n, sources, S, float(delta), 0, D, q, bits, stats)
    del q, bits
    if int(lib.pb200_sssp_overflowed(_native.stream_ptr())):
        exact_rerun()
    return D, stats


# ---- cell-decomposed search (sssp_cells.cu) ----
CELLS_MIN_N = 65536           # below this the per-source chain is short enough
CELL_TARGET = 4096            # vertices per cell aimed at (cells = n / CELL_TARGET seeds)
CELL_CAP = 16384              # larger cells are dissolved (pb200_cells_max_cell: bitmap of the in-cell kernel)
CELL_ALPHA = 1.0              # a cell is dissolved when boundary^2 > CELL_ALPHA * its arcs
CELL_UNIT_GROW = False        # grow the cells breadth-first (every arc counts 1) instead of by arc weight
CELL_LOCAL_DELTA = 1.0        # bucket width of the in-cell search, in units of the global one
_CELLINFO = np.dtype([("voff", "<i4"), ("size", "<i4"), ("nb", "<i4"), ("ovoff", "<i4"), ("kept", "<i4"),
                      ("pad", "<i4"), ("toff", "<i8")])
LAST_CELLS: dict = {}


def sssp_cells(g: dv.DeviceCSR, sources: torch.Tensor, delta: Optional[float] = None, want_stats: bool = False,
               cell_target: Optional[int] = None):
    """Shortest paths through the cell decomposition of sssp_cells.cu.  Returns (D [S, n] with its COLUMNS IN CELL
    ORDER, order int32 [n] with order[p] = vertex of g at column p, stats) or None when the graph does not lend
    itself to it (boundaries too large: not a thin graph) -- the caller then runs sssp_per_source."""
    n, S = g.n, int(sources.numel())
    lib = _native.load()
    if delta is None:
        delta = _default_delta(g)
    target = int(cell_target or CELL_TARGET)
    P = int(max(1, min(2048, round(n / target))))
    info = {"cells": P}
    with stage("sssp_partition"):
        seeds_h = np.unique(((np.arange(P) + 0.5) * n / P).astype(np.int32))
        P = len(seeds_h)
        seeds = dv.to_device(seeds_h)
        key = torch.full((n,), -1, dtype=I64, device=g.indptr.device)
        qcap = int(max(4096, 8 * (n // P + 1)))
        q = empty((P * 2 * qcap,), I32)
        call("pb200_cells_partition", g.indptr, g.indices, g.data, n, seeds, P, float(delta) / 2.5, int(CELL_UNIT_GROW),
             key, q, qcap)
        nbytes = int(lib.pb200_cells_scratch_bytes(n))
        scratch = empty((nbytes,), torch.uint8)
        cell = empty((n,), I32)
        counts = zeros((2 * (P + 1),), I32)
        size_d, nb_d = counts[:P + 1], counts[P + 1:]
        arcs_d = zeros((P + 1,), I64)
        order, inv = empty((n,), I32), empty((n,), I32)
        call("pb200_cells_order", g.indptr, g.indices, n, key, P, cell, size_d, nb_d, arcs_d, order, inv, scratch,
             nbytes)
        del q
        overflow = int(lib.pb200_cells_overflowed(_native.stream_ptr()))
        if _dist.world()[1] > 1:
            # the cells grow by first-come claims (within a margin), so two runs -- two ranks -- need not agree on
            # them; every rank must use the same cells (its distance rows share their column order with the other
            # ranks' in the reductions that follow): rank 0's are taken
            flag = torch.tensor([overflow], dtype=I32, device=cell.device)
            for t in (flag, cell, counts, arcs_d, order, inv):
                _dist.broadcast_(t)
            overflow = int(flag.item())
        if overflow:
            return None
        counts_h = counts.cpu().numpy().astype(np.int64)
        size_h, nb_h = counts_h[:P + 1], counts_h[P + 1:]
        arcs_h = arcs_d.cpu().numpy()
    kept = (np.arange(P + 1) < P) & (size_h > 0) & (size_h <= CELL_CAP) & (nb_h * nb_h <= CELL_ALPHA * arcs_h)
    ovn = np.where(kept, nb_h, size_h)
    n_ov = int(ovn.sum())
    est_arcs = int((nb_h[kept] ** 2).sum() + arcs_h[~kept].sum())
    info.update(kept=int(kept.sum()), overlay_vertices=n_ov, boundary=int(nb_h[kept].sum()),
                dissolved_vertices=int(size_h[~kept].sum()), unreached=int(size_h[P]))
    LAST_CELLS.clear(); LAST_CELLS.update(info)
    if n_ov > 0.5 * n or est_arcs > 0.5 * g.nnz:
        return None
    voff = np.concatenate([[0], np.cumsum(size_h)[:-1]])
    ovoff = np.concatenate([[0], np.cumsum(ovn)[:-1]])
    tsz = np.where(kept, nb_h * size_h, 0)
    toff = np.concatenate([[0], np.cumsum(tsz)[:-1]])
    t_total = int(tsz.sum())
    # sources: position in cell order, cell, and whether they need a table row of their own
    src_pos = inv[sources.long()].contiguous()
    src_pos_h = src_pos.cpu().numpy().astype(np.int64)
    src_cell = np.searchsorted(voff, src_pos_h, side="right") - 1
    src_local = src_pos_h - voff[src_cell]
    src_interior = kept[src_cell] & (src_local >= nb_h[src_cell])
    src_rows = np.where(src_interior, size_h[src_cell], 0)
    src_toff_h = np.where(src_interior, t_total + np.concatenate([[0], np.cumsum(src_rows)[:-1]]), -1).astype(np.int64)
    t_all = t_total + int(src_rows.sum())
    ci = np.zeros(P + 1, dtype=_CELLINFO)
    ci["voff"], ci["size"], ci["nb"], ci["ovoff"], ci["kept"], ci["toff"] = voff, size_h, nb_h, ovoff, kept, toff
    cells_d = dv.to_device(ci.view(np.uint8))
    # work items of the in-cell kernel: every (kept cell, boundary vertex) and every interior source
    kc = np.flatnonzero(kept & (nb_h > 0))
    it_cell = np.repeat(kc, nb_h[kc])
    it_src = np.arange(len(it_cell)) - np.repeat(np.cumsum(nb_h[kc]) - nb_h[kc], nb_h[kc])
    it_out = toff[it_cell] + it_src * size_h[it_cell]
    kc_items = it_cell
    sj = np.flatnonzero(src_interior)
    it_cell = np.concatenate([it_cell, src_cell[sj]])
    it_src = np.concatenate([it_src, src_local[sj]])
    it_out = np.concatenate([it_out, src_toff_h[sj]])
    T = empty((max(1, t_all),), F64)
    with stage("sssp_tables"):
        indptr2, indices2, wts2 = empty((n + 1,), I32), empty((g.nnz,), I32), empty((g.nnz,), F64)
        cell2, nint = empty((n,), I32), empty((n,), I32)
        call("pb200_cells_relabel", g.indptr, g.indices, g.data, n, order, inv, cell, indptr2, indices2, wts2, cell2,
             nint, scratch, nbytes)
        # the tables do not depend on the sources: with several ranks every rank computes the cells of a contiguous
        # range (balanced by boundary vertices x arcs) and the ranges are exchanged; the rows of a rank's own interior
        # sources are always its own
        rank, world = _dist.world()
        n_tab = len(kc_items)
        mine = np.ones(len(it_cell), dtype=bool)
        cut_cells = None
        if world > 1 and n_tab > 0:
            work = np.cumsum(nb_h[kc] * arcs_h[kc])
            cut_cells = np.concatenate([[0], np.searchsorted(work, work[-1] * np.arange(1, world) / world, side="right"),
                                        [len(kc)]])
            own_cells = np.zeros(P + 1, dtype=bool)
            own_cells[kc[cut_cells[rank]:cut_cells[rank + 1]]] = True
            mine[:n_tab] = own_cells[it_cell[:n_tab]]
        sel = np.flatnonzero(mine)
        sel = sel[np.argsort(it_cell[sel], kind="stable")]   # items of a cell next to each other: its arcs stay in L2 / L1
        call("pb200_cells_local_sssp", indptr2, nint, indices2, wts2, cells_d, dv.to_device(it_cell[sel].astype(np.int32)),
             dv.to_device(it_src[sel].astype(np.int32)), dv.to_device(it_out[sel].astype(np.int64)), int(sel.size),
             float(delta) * CELL_LOCAL_DELTA, T)
        if cut_cells is not None:
            for r in range(world):
                a, b = cut_cells[r], cut_cells[r + 1]
                if b > a:
                    lo = int(toff[kc[a]])
                    hi = int(toff[kc[b - 1]] + tsz[kc[b - 1]])
                    torch.distributed.broadcast(T[lo:hi], src=r)
    with stage("sssp_overlay"):
        src_toff = dv.to_device(src_toff_h)
        pos_ov, ov_pos = empty((n,), I32), empty((max(1, n_ov),), I32)
        rows = n_ov + S
        deg = empty((rows + 1,), I32)
        ov_indptr = empty((rows + 1,), I32)
        src_pos32 = src_pos.to(I32).contiguous()
        call("pb200_cells_overlay_count", indptr2, nint, n, cells_d, cell2, n_ov, src_pos32, S, pos_ov, ov_pos, deg)
        call("pb200_cells_overlay_scan", deg, rows + 1, ov_indptr, scratch, nbytes)
        nnz_ov = int(ov_indptr[rows].item())
        ov_indices, ov_wts = empty((max(1, nnz_ov),), I32), empty((max(1, nnz_ov),), F64)
        call("pb200_cells_overlay_fill", indptr2, nint, indices2, wts2, cells_d, cell2, pos_ov, ov_pos, n_ov, src_pos32,
             src_toff, S, T, ov_indptr, ov_indices, ov_wts)
        ov = dv.DeviceCSR(ov_indptr, ov_indices, ov_wts, rows)
        ov_sources = torch.arange(n_ov, n_ov + S, dtype=I64, device=g.indptr.device)
    Dov, stats = sssp_per_source(ov, ov_sources, delta, want_stats)
    with stage("sssp_expand"):
        ks = np.flatnonzero(kept)
        ntile = (size_h[ks] + 255) // 256
        t_cell = np.repeat(ks, ntile)
        t_col = (np.arange(len(t_cell)) - np.repeat(np.cumsum(ntile) - ntile, ntile)) * 256
        tiles = dv.to_device(np.stack([t_cell, t_col], 1).astype(np.int32)) if len(t_cell) else None
        ds = np.flatnonzero(~kept & (size_h > 0))
        dis = np.concatenate([np.arange(voff[c], voff[c] + size_h[c]) for c in ds]) if len(ds) else np.zeros(0, np.int64)
        dis_pos = dv.to_device(dis.astype(np.int32)) if len(dis) else None
        D = empty((S, n), F64)
        call("pb200_cells_expand", cells_d, cell2, tiles, int(len(t_cell)), dis_pos, int(len(dis)), pos_ov, src_pos32,
             src_toff, T, Dov, rows, S, D, n)
    info.update(overlay_arcs=nnz_ov, table_doubles=t_all, items=int(len(it_cell)))
    LAST_CELLS.update(info)
    return D, order, stats


def sssp(g: dv.DeviceCSR, sources: torch.Tensor, delta: Optional[float] = None, want_stats: bool = False,
         keep_cell_order: bool = False):
    """D[S, n] shortest-path distances from each source (unreachable = +inf).  Large graphs go through the cell
    decomposition (sssp_cells), falling back to the per-source kernel when it does not apply.  Returns (D, stats)
    with the columns in g's vertex order, or -- keep_cell_order=True -- (D, stats, order, inv) with column p holding
    vertex order[p] and inv[order[p]] = p (both None = g's own order): the caller then saves the pass over D that
    un-permutes it."""
    out = None
    if g.n >= CELLS_MIN_N and SSSP_MODE != "per_source":
        out = sssp_cells(g, sources, delta, want_stats)
    if out is None:
        D, stats = sssp_per_source(g, sources, delta, want_stats)
        return (D, stats, None, None) if keep_cell_order else (D, stats)
    D, order, stats = out
    if keep_cell_order:
        inv = empty((g.n,), I32)
        inv[order.long()] = torch.arange(g.n, dtype=I32, device=order.device)
        return D, stats, order, inv
    Dt = empty((g.n, D.shape[0]), F64)                  # un-permute the columns: rows of the transpose
    call("pb200_unpermute_rows_f64", D.t().contiguous(), g.n, D.shape[0], order, Dt)
    return Dt.t().contiguous(), stats


def connect_graph(idx: torch.Tensor, dist: torch.Tensor, data: torch.Tensor, start: int):
    """core.py:806-842 on the device.  Returns the undirected graph; when some cells are
    unreachable from `start` the reference's loop is followed literally: farthest reachable
    cell -> its nearest (sklearn euclidean) unreachable cell, one arc per iteration."""
    g = undirected_graph(idx, dist)
    n_unreach, label = unreachable_count(g, start)
    if n_unreach == 0:
        return g
    warnings.warn(
        "Some of the cells were unreachable. Consider increasing the k for \n \
            nearest neighbor graph construction.")
    n = g.n
    src = torch.tensor([start], dtype=I64, device=idx.device)
    pd = empty((n,), F64)
    while n_unreach > 0:
        D, _ = sssp(g, src)
        d0 = D[0]
        reach = torch.isfinite(d0)
        far = int(torch.argmax(torch.where(reach, d0, torch.full_like(d0, -1.0))).item())
        call("pb200_point_dists_expanded", data, n, data.shape[1], far, pd)
        cand = torch.where(reach, torch.full_like(pd, float("inf")), pd)
        j = int(torch.argmin(cand).item())
        w = pd[j:j + 1]
        # widen the arc table by one column: self everywhere (dropped) except far -> j
        extra_i = torch.arange(n, dtype=I32, device=idx.device).unsqueeze(1)
        extra_d = zeros((n, 1), F64)
        extra_i[far, 0] = j
        extra_d[far, 0] = w[0]
        idx = torch.cat([idx, extra_i], dim=1).contiguous()
        dist = torch.cat([dist, extra_d], dim=1).contiguous()
        g = undirected_graph(idx, dist)
        n_unreach, label = unreachable_count(g, start)
    return g


# ---------------------------------------------------------------------------
# refinement
# ---------------------------------------------------------------------------
def _sum_pow(x: torch.Tensor, shift: float, power: int) -> float:
    cnt = x.numel()
    partial = empty((cnt // 4096 + 2,), F64)
    out = empty((1,), F64)
    call("pb200_sum_pow_f64", x, cnt, float(shift), power, partial, out)
    return out


def bandwidth(D: torch.Tensor, total_count: int) -> float:
    """sdv = std(D) * 1.06 * count^(-1/5) (core.py:368); D may be this rank's row shard."""
    s1 = _dist.allreduce_sum_(_sum_pow(D, 0.0, 1))
    mean = float(s1.item()) / total_count
    s2 = _dist.allreduce_sum_(_sum_pow(D, mean, 2))
    std = math.sqrt(float(s2.item()) / total_count)
    return std * 1.06 * total_count ** (-1 / 5)


def pearson(x: torch.Tensor, y: torch.Tensor) -> float:
    n = x.numel()
    partial = empty((3 * (n // 4096 + 2),), F64)
    out = empty((3,), F64)
    mx = float(_sum_pow(x, 0.0, 1).item()) / n
    my = float(_sum_pow(y, 0.0, 1).item()) / n
    call("pb200_pearson_sums", x, y, n, mx, my, partial, out)
    sxy, sxx, syy = out.cpu().tolist()
    return sxy / math.sqrt(sxx * syy)


def refine_pseudotime(D: torch.Tensor, row0: int, n_rows_total: int, wp_cells: torch.Tensor, start_row: int,
                      max_iterations: int, verbose: bool = True):
    """core.py:368-397.  D holds waypoint rows row0..row0+S-1 of the full nW x N matrix
    (all of it on one GPU).  Returns (pseudotime[N] in [0,1], colsum[N], sdv)."""
    S, n = D.shape
    sdv = bandwidth(D, n_rows_total * n)
    colsum = empty((n,), F64)
    with stage("refine"):
        call("pb200_waypoint_colsum", D, S, n, sdv, colsum)
        _dist.allreduce_sum_(colsum)
        # initial pseudotime = distance row of the start waypoint
        if row0 <= start_row < row0 + S:
            pt = D[start_row - row0].clone()
        else:
            pt = zeros((n,), F64)
        _dist.allreduce_sum_(pt) if _dist.world()[1] > 1 else None
        plus_row = start_row - row0 if row0 <= start_row < row0 + S else -1
        my_cells = wp_cells[row0:row0 + S].contiguous()
        it, converged = 1, False
        while not converged and it < max_iterations:
            t_mask = gather_rows(pt, my_cells)
            t_add = t_mask.clone()
            if plus_row >= 0:
                t_add[plus_row] = 0.0
            new = empty((n,), F64)
            call("pb200_refine_step", D, S, n, sdv, colsum, pt, t_mask, t_add, plus_row, new)
            _dist.allreduce_sum_(new)
            corr = pearson(pt, new)
            if verbose:
                print("Correlation at iteration %d: %.4f" % (it, corr))
            converged = corr > 0.9999
            pt = new
            it += 1
        mn, _, mx, _ = col_minmax(pt.view(n, 1))
        lo, hi = float(mn.item()), float(mx.item())
        call("pb200_shift_scale", pt, n, lo, hi - lo)
    return pt, colsum, sdv


def project_fates(D: torch.Tensor, sdv: float, colsum: torch.Tensor, B: torch.Tensor):
    """fates[N, nT] = W_norm^T B and the per-cell entropy (core.py:225-230)."""
    S, n = D.shape
    nt = B.shape[1]
    fates = zeros((n, nt), F64)
    ent = empty((n,), F64)
    with stage("project"):
        if nt > 0:
            call("pb200_project_fates", D, S, n, sdv, colsum, B.contiguous(), nt, fates)
            _dist.allreduce_sum_(fates)
        call("pb200_row_entropy", fates, n, nt, ent)
    return fates, ent


# ---------------------------------------------------------------------------
# Markov chain on the waypoints and absorption probabilities
# ---------------------------------------------------------------------------
def markov_chain_dense(wp_data: torch.Tensor, knn: int, pt_wp: torch.Tensor) -> torch.Tensor:
    """Dense row-stochastic T[S, S] of core.py:523-563."""
    S = wp_data.shape[0]
    if knn > S:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {knn}, n_samples_fit = {S}")
    ind, dist, _ = dv.knn_like_sklearn(wp_data.contiguous(), knn)
    ka = int(min(int(math.floor(knn / 3)) - 1, 30))
    if ka < 0:
        ka += knn                      # numpy's negative index (knn < 3)
    T = zeros((S, S), F64)
    aff = empty((S * knn,), F64)
    rowsum = empty((S,), F64)
    with stage("markov_chain"):
        call("pb200_markov_chain_dense", ind, dist, S, knn, ka, pt_wp.contiguous(), aff, rowsum, T)
    return T


def absorption_probabilities(T: torch.Tensor, abs_states: np.ndarray):
    """core.py:699-765: B[S, nA] in waypoint order (transient rows solved, absorbing rows = identity)."""
    S = T.shape[0]
    abs_states = np.asarray(abs_states, dtype=np.int64)
    trans = np.setdiff1d(np.arange(S), abs_states)
    na, nt = len(abs_states), len(trans)
    B = zeros((S, na), F64)
    if na == 0:
        return B
    abs_d = dv.to_device(abs_states.astype(np.int32))
    B[torch.from_numpy(abs_states).to(B.device), torch.arange(na, device=B.device)] = 1.0
    if nt == 0:
        return B
    trans_d = dv.to_device(trans.astype(np.int32))
    A = empty((nt, nt), F64)
    R = empty((nt, na), F64)
    singular = empty((1,), I32)
    with stage("absorption_solve"):
        call("pb200_absorb_system", T, S, trans_d, nt, abs_d, na, A, R)
        call("pb200_dense_solve_f64", A, R, nt, na, singular)
    if int(singular.item()) != 0:
        raise RuntimeError("absorption system (I - Q) is singular")
    R.clamp_(min=0.0)                       # core.py:753
    B[torch.from_numpy(trans).to(B.device)] = R
    return B


def stationary_ranks(T: torch.Tensor, iters: int = 20000, tol: float = 1e-13) -> np.ndarray:
    """|leading left eigenvector| of T (core.py:599-600) by power iteration on the device,
    started from the uniform vector (the reference lets ARPACK pick a random start)."""
    S = T.shape[0]
    x = torch.full((S,), 1.0 / S, dtype=F64, device=T.device)
    y = empty((S,), F64)
    for it in range(iters):
        call("pb200_dense_matvec_t", T, S, x, y)
        nrm = float(_sum_pow(y, 0.0, 2).item()) ** 0.5
        if nrm == 0.0:
            break
        call("pb200_shift_scale", y, S, 0.0, nrm)
        if it % 50 == 49:
            diff = float(_sum_pow(y - x, 0.0, 2).item()) ** 0.5
            x, y = y, x
            if diff < tol:
                break
        else:
            x, y = y, x
    return np.abs(x.cpu().numpy())
