"""AnnData handling without a hard dependency on anndata (not installed in the build image).

``is_anndata`` accepts the real ``anndata.AnnData`` when it is importable and any
object exposing the attributes the hot path touches (``obsm``, ``obsp``, ``uns``,
``obs``, ``obs_names``).  ``AnnDataLike`` is the minimal stand-in the tests and the
bench use.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

try:  # pragma: no cover - not installed in this image
    from anndata import AnnData as _RealAnnData
except Exception:  # noqa: BLE001
    _RealAnnData = None

_ATTRS = ("obsm", "obsp", "uns", "obs", "obs_names")


def is_anndata(obj) -> bool:
    if _RealAnnData is not None and isinstance(obj, _RealAnnData):
        return True
    if isinstance(obj, (pd.DataFrame, np.ndarray, dict, str)):
        return False
    return all(hasattr(obj, a) for a in _ATTRS)


class AnnDataLike:
    """Dict-backed stand-in: X, obs (DataFrame), obs_names, var_names, obsm, obsp, uns, layers."""

    def __init__(self, X=None, obs_names=None, var_names=None):
        self.X = X
        n = 0 if X is None else X.shape[0]
        if obs_names is None:
            obs_names = pd.Index([str(i) for i in range(n)])
        self.obs = pd.DataFrame(index=pd.Index(obs_names))
        g = 0 if X is None or len(X.shape) < 2 else X.shape[1]
        self.var_names = pd.Index(var_names) if var_names is not None else pd.Index([str(i) for i in range(g)])
        self.obsm, self.obsp, self.uns, self.layers = {}, {}, {}, {}

    @property
    def obs_names(self) -> pd.Index:
        return self.obs.index

    @property
    def n_obs(self) -> int:
        return len(self.obs.index)
