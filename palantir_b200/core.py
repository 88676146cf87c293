"""Pseudotime and fate-probability API of Palantir on B200: same signature, defaults,
prints, warnings, return type and AnnData keys as ``palantir.core.run_palantir``
(reference: src/palantir/core.py:61-249), with the arithmetic on the CUDA library.

Host side (this file): cell-label bookkeeping with pandas exactly as the reference
does it (label unions are sorted by pandas, fate columns follow waypoint order), the
MT19937 draws of the first waypoint per component, and assembling the pandas outputs.
Device side (``_device_core``): min-max scaling, boundary cells, max-min sampling,
exact kNN graph, reachability, batched multi-source shortest paths, pseudotime
refinement, waypoint Markov chain, dense FP64 absorption solve, projection + entropy.
"""
from __future__ import annotations

import time
import warnings
from typing import Optional, Tuple

import numpy as np
import pandas as pd
import torch

from . import _devcache
from . import _device as dv
from . import _device_core as dc
from . import _dist, config
from .compat import is_anndata
from .validation import normalize_cell_identifiers

__all__ = ["run_palantir", "identify_terminal_states"]

LAST_INFO: dict = {}


def _positions(index: pd.Index, labels, what: str) -> np.ndarray:
    """Positions of ``labels`` in ``index``; a label that is not a cell of the data raises KeyError, as the
    reference's ``data.loc[labels]`` does (core.py:209, 355, 501) -- never a -1 that would index before the
    first row on the device."""
    pos = index.get_indexer(pd.Index(labels))
    if (pos < 0).any():
        missing = [l for l, p_ in zip(labels, pos) if p_ < 0]
        raise KeyError(f"{what} not found in the data index: {missing[:10]}" + (" ..." if len(missing) > 10 else ""))
    return np.asarray(pos, dtype=np.int64)


def _get_joblib_backend():
    """Kept for API compatibility (core.py:31-58); the B200 path starts no joblib workers."""
    return config.JOBLIB_BACKEND


def _max_min_sampling(data: pd.DataFrame, num_waypoints: int, seed: Optional[int] = None,
                      _device_data: Optional[torch.Tensor] = None) -> pd.Index:
    """Max-min waypoint sampling (core.py:252-312).  The first waypoint of every column is
    drawn with the reference's own RNG calls (np.random.seed(seed) once, then one
    np.random.randint(N) per column); the farthest-point iterations run on the device."""
    n, m = data.shape
    min_waypoints = max(3, m)
    if num_waypoints < min_waypoints:
        warnings.warn(
            f"num_waypoints={num_waypoints} is too small for {m} components; "
            f"using {min_waypoints} instead for stable sampling.", UserWarning)
        num_waypoints = min_waypoints
    no_iterations = int(num_waypoints / m)
    if seed is not None:
        np.random.seed(seed)
    first = np.array([np.random.randint(n) for _ in range(m)], dtype=np.int64)
    dev = _device_data if _device_data is not None else dv.to_device(np.ascontiguousarray(data.values, dtype=np.float64))
    picks = dc.maxmin_sampling(dev, first, no_iterations)          # [m, iters]
    return data.index[picks.reshape(-1)].unique()


def _boundary_positions(dev: torch.Tensor) -> np.ndarray:
    """Positions of the per-column arg-max and arg-min cells (core.py:186)."""
    _, imin, _, imax = dc.col_minmax(dev)
    return np.concatenate([imax.cpu().numpy(), imin.cpu().numpy()])


def _compute_pseudotime_device(dev: torch.Tensor, start_pos: int, knn: int, wp_pos: np.ndarray,
                               max_iterations: int):
    """core.py:353-401 on positions.  When sklearn would pick its kd-tree (<= 15 columns) the
    stage runs on cells ordered along the Morton curve of the leading components (box-pruned exact
    kNN, L2-friendly shortest paths).  Returns (pseudotime [N] in input order, D shard [rows, N] and colsum [N] in
    COLUMN order, row0, sdv, perm) with perm = None when column order == input order, else
    column i = input cell perm[i] (Morton work order composed with the cell order of the shortest-path stage)."""
    print("Shortest path distances using {}-nearest neighbor graph...".format(knn))
    t0 = time.time()
    n, m = dev.shape
    if knn > n:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {knn}, n_samples_fit = {n}")
    wp_pos = np.asarray(wp_pos, dtype=np.int64)
    if not dv.sklearn_uses_brute(n, m, knn):
        perm, inv, work = dc.morton_order(dev)
        idx, dist, kstats = dc.knn_table_boxed(work, knn)
        inv_h = inv.cpu().numpy().astype(np.int64)
        start_w, wp_w = int(inv_h[start_pos]), inv_h[wp_pos]
    else:
        perm, work = None, dev
        idx, dist, kstats = dc.knn_table(dev, knn)
        start_w, wp_w = int(start_pos), wp_pos
    g = dc.connect_graph(idx, dist, work, start_w)
    n_wp = len(wp_w)
    rank, size = _dist.world()
    bounds = _dist.shard_bounds(n_wp, size)
    row0, row1 = bounds[rank], bounds[rank + 1]
    sources = dv.to_device(np.ascontiguousarray(wp_w[row0:row1], dtype=np.int64))
    D, stats, order, order_inv = dc.sssp(g, sources, want_stats=dv.PROFILE is not None, keep_cell_order=True)
    if dv.PROFILE is not None:
        torch.cuda.synchronize()
        if "scanned" in kstats:
            kstats = dict(kstats, scanned=int(kstats["scanned"].item()))
    kstats = {k: v for k, v in kstats.items() if not isinstance(v, torch.Tensor)}
    LAST_INFO["sssp"] = {
        "sources": int(row1 - row0), "arcs": int(g.nnz), "knn": kstats,
        "cells": dict(dc.LAST_CELLS) if order is not None else None,
        "rounds_mean": float(stats[:, 0].double().mean().item()) if stats is not None and stats.numel() else None,
        "relax_mean": float(stats[:, 1].double().mean().item()) if stats is not None and stats.numel() else None,
        "advances_mean": float(stats[:, 2].double().mean().item()) if stats is not None and stats.numel() else None,
    }
    print("Time for shortest paths: {} minutes".format((time.time() - t0) / 60))
    print("Iteratively refining the pseudotime...")
    start_row = int(np.flatnonzero(wp_w == start_w)[0])
    wp_cells = dv.to_device(np.ascontiguousarray(wp_w, dtype=np.int64))
    if order is not None:
        # the cell-decomposed search leaves the columns of D in its own cell order: column p = work vertex order[p]
        wp_cells = order_inv[wp_cells].long().contiguous()
        perm = order if perm is None else perm[order.long()].contiguous()
    pt_w, colsum, sdv = dc.refine_pseudotime(D, row0, n_wp, wp_cells, start_row, max_iterations)
    pt = pt_w if perm is None else dc.unpermute_rows(pt_w, perm)
    return pt, D, row0, colsum, sdv, perm


def _compute_pseudotime(data: pd.DataFrame, start_cell, knn: int, waypoints: pd.Index, n_jobs: int,
                        max_iterations: int = 25) -> Tuple[pd.Series, pd.DataFrame]:
    """Reference-shaped helper (core.py:315-401): returns (pseudotime Series, W DataFrame).
    W (nW x N) is materialised on the host only here, for callers and tests that want it;
    run_palantir never forms it."""
    if knn is None or int(knn) <= 0:
        raise ValueError("knn must be a positive integer")
    dev = dv.to_device(np.ascontiguousarray(data.values, dtype=np.float64))
    start_pos = int(_positions(data.index, [start_cell], "start cell")[0])
    wp_pos = _positions(data.index, waypoints, "waypoints")
    pt, D, row0, colsum, sdv, perm = _compute_pseudotime_device(dev, start_pos, int(knn), wp_pos, max_iterations)
    if _dist.world()[1] > 1:
        raise NotImplementedError("_compute_pseudotime returns the dense W and is single-GPU only")
    Dh = D.cpu().numpy()
    W = np.exp(-0.5 * np.power(Dh / sdv, 2)) / colsum.cpu().numpy()[None, :]
    if perm is not None:                      # columns back to the input order
        Wo = np.empty_like(W)
        Wo[:, perm.cpu().numpy()] = W
        W = Wo
    return (pd.Series(pt.cpu().numpy(), index=data.index),
            pd.DataFrame(W, index=waypoints, columns=data.index))


def _terminal_states_from_markov_chain(T: torch.Tensor, wp_dev: torch.Tensor, wp_labels: pd.Index,
                                       pt_wp: np.ndarray) -> np.ndarray:
    """core.py:590-649 (labels out).  The stationary vector comes from a device power
    iteration with a fixed start (the reference's ARPACK start is random); the MAD cut-off
    and the component bookkeeping over the few selected waypoints are host arithmetic."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.stats import norm

    print("Identification of terminal states...")
    bpos = np.unique(_boundary_positions(wp_dev))
    ranks = dc.stationary_ranks(T)
    med = np.median(ranks)
    cutoff = norm.ppf(0.9999, loc=med, scale=np.median(np.abs(ranks - med)))
    sel = np.flatnonzero(ranks > cutoff)
    if sel.size == 0:
        return np.array([])
    sel_d = torch.from_numpy(sel).to(T.device)
    sub = T[sel_d][:, sel_d].cpu().numpy()
    n_comp, lab = connected_components(csr_matrix(sub), directed=False)
    cand = []
    for c in range(n_comp):
        members = sel[lab == c]
        cand.append(int(members[int(np.argmax(pt_wp[members]))]))
    wp_host = wp_dev.cpu().numpy()
    # 2m boundary cells x a handful of candidates: label bookkeeping on the host
    diff = wp_host[bpos, :][:, None, :] - wp_host[cand, :][None, :, :]
    dd = np.sqrt((diff * diff).sum(axis=2))
    return np.unique(np.asarray(wp_labels)[bpos[np.argmin(dd, axis=0)]])


def _differentiation_entropy_device(wp_dev: torch.Tensor, wp_labels: pd.Index, terminal_states, knn: int,
                                    pt_wp_dev: torch.Tensor):
    """core.py:652-765: returns (terminal labels in waypoint order, B[S, nT] device)."""
    print("Markov chain construction...")
    T = dc.markov_chain_dense(wp_dev, knn, pt_wp_dev)
    if terminal_states is None:
        terminal_states = _terminal_states_from_markov_chain(T, wp_dev, wp_labels, pt_wp_dev.cpu().numpy())
    abs_pos = np.where(wp_labels.isin(terminal_states))[0]
    if len(abs_pos) == 0:
        return wp_labels[abs_pos], dv.zeros((len(wp_labels), 0), dv.F64)
    print("Computing fundamental matrix and absorption probabilities...")
    B = dc.absorption_probabilities(T, abs_pos)
    return wp_labels[abs_pos], B


def run_palantir(data, early_cell, terminal_states=None, knn: int = 30, num_waypoints: int = 1200,
                 n_jobs: int = -1, scale_components: bool = True, use_early_cell_as_start: bool = False,
                 max_iterations: int = 25, eigvec_key: str = "DM_EigenVectors_multiscaled",
                 pseudo_time_key: str = "palantir_pseudotime", entropy_key: str = "palantir_entropy",
                 fate_prob_key: str = "palantir_fate_probabilities", save_as_df: bool = None,
                 waypoints_key: str = "palantir_waypoints", seed: int = 20):
    """Palantir pseudotime, entropy and fate probabilities (reference core.py:61-249).

    Returns a PResults in every case; for an AnnData input it also writes
    obs[pseudo_time_key], obs[entropy_key], uns[waypoints_key] and
    obsm[fate_prob_key] (+ uns[fate_prob_key + "_columns"] when not saved as DataFrame).
    ``n_jobs`` is accepted and ignored."""
    from .presults import PResults

    if save_as_df is None:
        save_as_df = config.SAVE_AS_DF
    if knn is None or int(knn) <= 0:
        raise ValueError("knn must be a positive integer")
    knn = int(knn)

    obs_names = data.obs_names if is_anndata(data) else data.index
    if terminal_states is not None:
        if isinstance(terminal_states, dict):
            terminal_states = pd.Series(terminal_states)
        terminal_dict, _, _ = normalize_cell_identifiers(
            terminal_states, obs_names, context="run_palantir terminal_states")
        if isinstance(terminal_states, pd.Series):
            terminal_states = pd.Series(terminal_dict)
        terminal_cells = list(terminal_dict.keys())
    else:
        terminal_cells = None

    if is_anndata(data):
        ms_data = pd.DataFrame(data.obsm[eigvec_key], index=data.obs_names)
    else:
        ms_data = data

    early_cell_dict, _, n_early_found = normalize_cell_identifiers(
        [early_cell], ms_data.index, context="run_palantir early_cell")
    if n_early_found == 0:
        raise ValueError(
            f"Early cell '{early_cell}' not found in data. "
            f"Please provide a valid cell identifier from data.obs_names (type: {type(ms_data.index[0]).__name__})")
    early_cell = list(early_cell_dict.keys())[0]

    index = ms_data.index
    n = len(index)
    ms_vals = ms_data.values
    ms_dev = _devcache.lookup(ms_vals)             # still on the device from determine_multiscale_space?
    if ms_dev is None or ms_dev.shape != ms_vals.shape or ms_dev.dtype != torch.float64:
        ms_dev = dv.to_device(ms_vals if ms_vals.dtype == np.float64 else ms_vals.astype(np.float64))
    if _dist.world()[1] > 1:
        # one process per GPU: every rank shards the SAME kNN rows / shortest-path sources, so all of them must
        # hold rank 0's matrix bit for bit (a last-bit or sign difference would mix rows of different graphs)
        ms_dev = _dist.broadcast_(ms_dev.contiguous())
    dev = dc.minmax_scale(ms_dev) if scale_components else ms_dev.clone()

    # boundary cells and the start cell (core.py:186-192)
    bpos = _boundary_positions(dev)
    dm_boundaries = pd.Index(set(index[bpos]))
    early_pos = int(index.get_loc(early_cell))
    b_list = index.get_indexer(dm_boundaries)
    pd_all = dv.empty((n,), dv.F64)
    dv.call("pb200_point_dists_expanded", dev, n, dev.shape[1], early_pos, pd_all)
    b_d = pd_all[torch.from_numpy(np.asarray(b_list)).to(pd_all.device)].cpu().numpy()
    start_cell = pd.Series(b_d, index=dm_boundaries).idxmin()
    if use_early_cell_as_start:
        start_cell = early_cell

    print("Sampling and flocking waypoints...")
    t0 = time.time()
    if isinstance(num_waypoints, (int, np.integer)):
        waypoints = _max_min_sampling_labels(index, ms_data.shape[1], int(num_waypoints), seed, dev)
    else:
        waypoints = num_waypoints
    waypoints = waypoints.union(dm_boundaries)
    if terminal_cells is not None:
        waypoints = waypoints.union(terminal_cells)
    waypoints = pd.Index(waypoints.difference([start_cell]).unique())
    waypoints = pd.Index([start_cell]).append(waypoints)
    print("Time for determining waypoints: {} minutes".format((time.time() - t0) / 60))

    print("Determining pseudotime...")
    start_pos = int(index.get_loc(start_cell))
    wp_pos = _positions(index, waypoints, "waypoints / terminal states")
    pt_dev, D, row0, colsum, sdv, perm = _compute_pseudotime_device(dev, start_pos, knn, wp_pos, max_iterations)

    print("Entropy and branch probabilities...")
    wp_pos_d = dv.to_device(np.ascontiguousarray(wp_pos, dtype=np.int64))
    wp_dev = dc.gather_rows(dev, wp_pos_d)
    pt_wp_dev = dc.gather_rows(pt_dev, wp_pos_d)
    term_labels, B = _differentiation_entropy_device(wp_dev, waypoints, terminal_cells, knn, pt_wp_dev)

    print("Project results to all cells...")
    B_local = B[row0:row0 + D.shape[0]].contiguous()
    fates_dev, ent_dev = dc.project_fates(D, sdv, colsum, B_local)
    if perm is not None:
        fates_dev, ent_dev = dc.unpermute_rows(fates_dev, perm), dc.unpermute_rows(ent_dev, perm)

    # PResults zeroes probabilities below 1 % (presults.py:34); the entropy above is of the unclipped rows
    if fates_dev.numel():
        dv.call("pb200_vec_transform", fates_dev, fates_dev.numel(), 4, 0.01)
    pt_h, fates_h, ent_h = dv.to_host(pt_dev, fates_dev, ent_dev)
    pseudotime = pd.Series(pt_h, index=index)
    branch_probs = pd.DataFrame(fates_h, index=index, columns=term_labels)
    if branch_probs.shape[1] == 0:
        ent = pd.Series(0.0, index=index)      # scipy entropy of an empty row is 0 in the reference's apply
    else:
        ent = pd.Series(ent_h, index=index)

    pr_res = PResults(pseudotime, ent, branch_probs, waypoints, _clipped=True)

    if is_anndata(data):
        data.obs[pseudo_time_key] = pseudotime
        data.obs[entropy_key] = ent
        data.uns[waypoints_key] = waypoints.values
        if isinstance(terminal_states, pd.Series):
            branch_probs.columns = terminal_states[branch_probs.columns]
        if save_as_df:
            data.obsm[fate_prob_key] = branch_probs
        else:
            data.obsm[fate_prob_key] = branch_probs.values
            data.uns[fate_prob_key + "_columns"] = branch_probs.columns.values
    return pr_res


def _max_min_sampling_labels(index: pd.Index, m: int, num_waypoints: int, seed, dev: torch.Tensor) -> pd.Index:
    """_max_min_sampling on an already-resident device matrix (labels from `index`)."""
    n = len(index)
    min_waypoints = max(3, m)
    if num_waypoints < min_waypoints:
        warnings.warn(
            f"num_waypoints={num_waypoints} is too small for {m} components; "
            f"using {min_waypoints} instead for stable sampling.", UserWarning)
        num_waypoints = min_waypoints
    no_iterations = int(num_waypoints / m)
    if seed is not None:
        np.random.seed(seed)
    first = np.array([np.random.randint(n) for _ in range(m)], dtype=np.int64)
    picks = dc.maxmin_sampling(dev, first, no_iterations)
    return index[picks.reshape(-1)].unique()


def identify_terminal_states(ms_data: pd.DataFrame, early_cell, knn: int = 30, num_waypoints: int = 1200,
                             n_jobs: int = -1, max_iterations: int = 25, seed: int = 20):
    """Terminal states from the waypoint Markov chain (reference core.py:404-491).
    Returns (terminal_states array of labels, excluded_boundaries Index)."""
    early_cell_dict, _, n_found = normalize_cell_identifiers(
        [early_cell], ms_data.index, context="identify_terminal_states early_cell")
    if n_found == 0:
        raise ValueError(
            f"Early cell '{early_cell}' not found in data. "
            f"Please provide a valid cell identifier from data.index (type: {type(ms_data.index[0]).__name__})")
    early_cell = list(early_cell_dict.keys())[0]
    index = ms_data.index
    n = len(index)
    dev = dc.minmax_scale(dv.to_device(np.ascontiguousarray(ms_data.values, dtype=np.float64)))
    bpos = _boundary_positions(dev)
    dm_boundaries = pd.Index(set(index[bpos]))
    pd_all = dv.empty((n,), dv.F64)
    dv.call("pb200_point_dists_expanded", dev, n, dev.shape[1], int(index.get_loc(early_cell)), pd_all)
    b_list = index.get_indexer(dm_boundaries)
    b_d = pd_all[torch.from_numpy(np.asarray(b_list)).to(pd_all.device)].cpu().numpy()
    start_cell = pd.Series(b_d, index=dm_boundaries).idxmin()
    waypoints = _max_min_sampling_labels(index, ms_data.shape[1], num_waypoints, seed, dev)
    waypoints = waypoints.union(dm_boundaries)
    waypoints = pd.Index(waypoints.difference([start_cell]).unique())
    waypoints = pd.Index([start_cell]).append(waypoints)
    wp_pos = _positions(index, waypoints, "waypoints")
    pt_dev = _compute_pseudotime_device(dev, int(index.get_loc(start_cell)), int(knn), wp_pos, max_iterations)[0]
    wp_pos_d = dv.to_device(np.ascontiguousarray(wp_pos, dtype=np.int64))
    wp_dev = dc.gather_rows(dev, wp_pos_d)
    pt_wp_dev = dc.gather_rows(pt_dev, wp_pos_d)
    print("Markov chain construction...")
    T = dc.markov_chain_dense(wp_dev, int(knn), pt_wp_dev)
    terminal_states = _terminal_states_from_markov_chain(T, wp_dev, waypoints, pt_wp_dev.cpu().numpy())
    wb = np.unique(_boundary_positions(wp_dev))
    wp_boundaries = pd.Index(set(waypoints[wb]))
    excluded = wp_boundaries.difference(terminal_states).difference([start_cell])
    return terminal_states, excluded
