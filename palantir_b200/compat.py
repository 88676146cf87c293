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
        self.var = pd.DataFrame(index=self.var_names)
        self.obsm, self.obsp, self.uns, self.layers, self.varm = {}, {}, {}, {}, {}

    @property
    def obs_names(self) -> pd.Index:
        return self.obs.index

    @property
    def n_obs(self) -> int:
        return len(self.obs.index)

    @property
    def n_vars(self) -> int:
        return len(self.var_names)

    # ---- on-disk round trip with the key layout an h5ad file keeps (anndata / h5py are not installed in the
    # build image; with the real anndata use ad.write_h5ad / anndata.read_h5ad) ----
    def write(self, path: str) -> None:
        """One .npz holding X, obs / var columns, and every obsm / obsp / uns / layers / varm entry under
        '<slot>/<key>'; DataFrames keep their columns ('<slot>/<key>@columns'), sparse matrices their CSR parts."""
        from scipy.sparse import issparse

        out = {"obs_names": np.asarray(self.obs_names, dtype=str), "var_names": np.asarray(self.var_names, dtype=str)}

        def put(name, v):
            if isinstance(v, pd.DataFrame):
                out[name] = v.values
                out[name + "@columns"] = np.asarray(v.columns, dtype=str)
            elif isinstance(v, pd.Series):
                out[name] = v.values
            elif issparse(v):
                c = v.tocsr()
                out[name + "@csr_data"], out[name + "@csr_indices"] = c.data, c.indices
                out[name + "@csr_indptr"], out[name + "@csr_shape"] = c.indptr, np.asarray(c.shape)
            elif isinstance(v, dict):
                for k2, v2 in v.items():
                    put(name + "." + k2, v2)
            else:
                out[name] = np.asarray(v)
            if name in out and out[name].dtype == object:
                out[name] = out[name].astype(str)           # labels: stored as fixed-width text, as h5ad stores strings

        if self.X is not None:
            put("X", self.X)
        for col in self.obs.columns:
            put("obs/" + col, self.obs[col])
        for col in self.var.columns:
            put("var/" + col, self.var[col])
        for slot in ("obsm", "obsp", "uns", "layers", "varm"):
            for k, v in getattr(self, slot).items():
                put(slot + "/" + k, v)
        np.savez(path, **out)

    @classmethod
    def read(cls, path: str) -> "AnnDataLike":
        from scipy.sparse import csr_matrix

        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz", allow_pickle=False)
        keys = list(z.keys())

        def get(name):
            if name + "@csr_data" in z:
                return csr_matrix((z[name + "@csr_data"], z[name + "@csr_indices"], z[name + "@csr_indptr"]),
                                  shape=tuple(z[name + "@csr_shape"]))
            return z[name]

        ad = cls(get("X") if ("X" in z or "X@csr_data" in z) else None, obs_names=z["obs_names"], var_names=z["var_names"])
        if ad.X is None:
            ad.obs = pd.DataFrame(index=pd.Index(z["obs_names"]))
        names = {k.split("@")[0] for k in keys}
        for name in sorted(names):
            if "/" not in name:
                continue
            slot, key = name.split("/", 1)
            val = get(name)
            if name + "@columns" in z:
                idx = ad.obs_names if slot == "obsm" else (ad.var_names if slot == "varm" else None)
                val = pd.DataFrame(val, index=idx, columns=z[name + "@columns"])
            if slot == "obs":
                ad.obs[key] = val
            elif slot == "var":
                ad.var[key] = val
            elif "." in key:                       # nested dict entry (uns['pca']['variance'] ...)
                top, sub = key.split(".", 1)
                getattr(ad, slot).setdefault(top, {})[sub] = val
            else:
                getattr(ad, slot)[key] = val
        return ad
