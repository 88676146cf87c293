"""Synthetic branching-trajectory input for the diffusion-map / fate-probability path.

This is the generator SURVEY.md §8(d) specifies (there is no dataset on the GPU
box and the reference's bundled marrow h5ad is absent from the checkout): a trunk
plus three branches of unequal length and mass in a 4-D latent space, rotated
into ``d`` PCA-like dimensions with isotropic noise.  Unequal branches keep the
diffusion eigenvalues separated so eigenvector parity is well defined.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

__all__ = ["synth_branching", "synth_names"]

_BRANCH_P = (0.20, 0.40, 0.25, 0.15)       # trunk, B0, B1, B2 mass
_BRANCH_LEN = (1.0, 0.7, 0.45)             # branch lengths


def synth_names(n: int) -> pd.Index:
    """Cell labels ``c0000000`` ... (string sort order == positional order)."""
    return pd.Index(np.char.add("c", np.char.zfill(np.arange(n).astype(str), 7)))


def synth_branching(n: int, d: int, seed: int = 0, noise: float = 0.01, dtype=np.float32):
    """Return ``(X, info)``.

    ``X`` is an ``n x d`` array (``dtype``), ``info`` a dict with the latent time,
    the component of every cell, the early cell position and one terminal cell
    position per branch (the maximum-time cell of that branch).
    """
    if d < 4:
        raise ValueError("d must be >= 4")
    rng = np.random.default_rng(seed)
    comp = rng.choice(4, size=n, p=_BRANCH_P)
    u = rng.random(n)
    lat = np.zeros((n, 4))
    time = np.empty(n)
    trunk = comp == 0
    lat[trunk, 0] = u[trunk]
    time[trunk] = u[trunk]
    for b in range(3):
        sel = comp == b + 1
        s = u[sel] * _BRANCH_LEN[b]
        lat[sel, 0] = 1.0 + 0.3 * s * (1 - b)
        lat[sel, 1 + b] = s
        time[sel] = 1.0 + s
    q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    x = lat @ q[:4, :] + noise * rng.normal(size=(n, d))
    x = np.ascontiguousarray(x.astype(dtype))
    trunk_idx = np.flatnonzero(trunk)
    early = int(trunk_idx[np.argmin(time[trunk_idx])]) if trunk_idx.size else int(np.argmin(time))
    terminals = {}
    for b in range(3):
        idx = np.flatnonzero(comp == b + 1)
        if idx.size:
            terminals[f"B{b}"] = int(idx[np.argmax(time[idx])])
    return x, {"time": time, "component": comp, "early": early, "terminals": terminals}
