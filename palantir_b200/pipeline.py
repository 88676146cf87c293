"""Device-resident run of the whole hot path (no host objects): the same stages and
kernels the public API drives (``utils.run_diffusion_maps`` ->
``utils.determine_multiscale_space`` -> ``core.run_palantir``), with the input matrix
already in HBM and every result left there.  ``bench.py`` times this for the
kernel-side throughput (``value``) and the public API for ``e2e``; positions stand in
for cell labels (zero-padded names sort like positions, so the waypoint order is the
one the labelled API produces).
"""
from __future__ import annotations

import contextlib
import io
from typing import Optional, Sequence

import numpy as np
import pandas as pd
import torch

from . import _device as dv
from . import _device_core as dc
from . import _dist, core, utils


def run_device(x_dev: torch.Tensor, early_pos: int, terminal_pos: Optional[Sequence[int]], knn: int = 30,
               n_components: int = 10, alpha: float = 0.0, num_waypoints: int = 1200, dm_seed: int = 0,
               seed: int = 20, max_iterations: int = 25, quiet: bool = True) -> dict:
    n = x_dev.shape[0]
    sink = io.StringIO()
    with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
        # ---- run_diffusion_maps ----
        with dv.stage("knn_total"):
            idx, dist, kstats = dv.knn_like_sklearn(x_dev, min(knn + 1, n), shard=True)
        order = kstats.pop("perm", None)
        kern = dv.adaptive_kernel(idx, dist, knn, alpha)
        kern.order = order
        tvals, lam, U = utils._diffusion_maps_device(kern, n_components, dm_seed)
        # ---- determine_multiscale_space (eigen-gap rule on 10 numbers, column scaling) ----
        lam_t = torch.from_numpy(np.ascontiguousarray(lam)).to(U.device)
        _dist.broadcast_(lam_t)
        lam = lam_t.cpu().numpy()
        gaps = lam[:-1] - lam[1:]
        n_eigs = int(np.argsort(gaps)[-1] + 1)
        if n_eigs < 3:
            n_eigs = int(np.argsort(gaps)[-2] + 1)
        if n_eigs < 3:
            n_eigs = 3
        ev = lam[1:n_eigs]
        ms = (U[:, 1:n_eigs] * torch.from_numpy(ev / (1 - ev)).to(U.device)).contiguous()
        _dist.broadcast_(ms)
        # ---- run_palantir ----
        dev = dc.minmax_scale(ms)
        bpos = np.unique(core._boundary_positions(dev))
        pd_all = dv.empty((n,), dv.F64)
        dv.call("pb200_point_dists_expanded", dev, n, dev.shape[1], int(early_pos), pd_all)
        bd = pd_all[torch.from_numpy(bpos).to(pd_all.device)].cpu().numpy()
        start_pos = int(bpos[int(np.argmin(bd))])
        positions = pd.RangeIndex(n)
        sampled = core._max_min_sampling_labels(positions, dev.shape[1], int(num_waypoints), seed, dev)
        wp = np.union1d(np.asarray(sampled, dtype=np.int64), bpos)
        if terminal_pos is not None:
            wp = np.union1d(wp, np.asarray(list(terminal_pos), dtype=np.int64))
        wp = np.concatenate([[start_pos], wp[wp != start_pos]]).astype(np.int64)
        pt, D, row0, colsum, sdv, perm = core._compute_pseudotime_device(dev, start_pos, knn, wp, max_iterations)
        wp_d = dv.to_device(wp)
        wp_dev = dc.gather_rows(dev, wp_d)
        pt_wp = dc.gather_rows(pt, wp_d)
        term_labels, B = core._differentiation_entropy_device(
            wp_dev, pd.Index(wp), None if terminal_pos is None else list(terminal_pos), knn, pt_wp)
        fates, ent = dc.project_fates(D, sdv, colsum, B[row0:row0 + D.shape[0]].contiguous())
        if perm is not None:
            fates, ent = dc.unpermute_rows(fates, perm), dc.unpermute_rows(ent, perm)
    return dict(kernel=kern, t_values=tvals, eigenvalues=lam, eigenvectors=U, multiscale=ms, pseudotime=pt,
                fate_probs=fates, entropy=ent, waypoints=wp, terminal=np.asarray(term_labels), knn=kstats,
                nnz=kern.nnz, D_rows=int(D.shape[0]))
