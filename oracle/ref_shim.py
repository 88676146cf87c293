"""Load the reference's own hot-path modules from /root/reference (this container only).

TEST INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` to pin the oracle and to
produce the fixtures under ``tests/golden/``.  ``/root/reference`` does not exist
on the GPU box, so nothing that runs there imports this file.

``import palantir`` from the reference fails here (matplotlib / scanpy / anndata /
mellon are not installed), so the five modules on the path are exec'd into a
private package ``_palantir_reference`` under stub ``matplotlib`` / ``anndata`` /
``scanpy`` modules.  One source-string patch is applied:
pandas >= 3 copy-on-write makes ``bp.values[...] = 1`` (core.py:714, :762) raise,
so the identity frame is built with ``np.eye`` instead (same values).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = "/root/reference/src/palantir"
_PKG = "_palantir_reference"


def available() -> bool:
    return os.path.isdir(REF_ROOT)


def load_reference(eig_tol=None):
    """Return ``(core, utils)`` modules of the reference.  ``eig_tol`` replaces the
    ARPACK ``tol=1e-4`` of utils.py:484 (the "tight" oracle variant)."""
    if not available():
        raise RuntimeError("reference checkout not present at " + REF_ROOT)
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "anndata", "scanpy")}
    mpl = types.ModuleType("matplotlib"); mpl.rcParams = {}
    an = types.ModuleType("anndata")

    class AnnData:  # isinstance() checks only; DataFrame code paths are used
        pass

    an.AnnData = AnnData
    sc = types.ModuleType("scanpy"); sc.AnnData = AnnData
    sys.modules.update(matplotlib=mpl, anndata=an, scanpy=sc)
    suffix = "" if eig_tol is None else f"_tol{eig_tol:g}".replace("-", "m")
    pkgname = _PKG + suffix
    pkg = types.ModuleType(pkgname); pkg.__path__ = [REF_ROOT]
    sys.modules[pkgname] = pkg

    def load(name, patches=()):
        path = f"{REF_ROOT}/{name}.py"
        src = open(path).read()
        for old, new in patches:
            assert old in src, (name, old)
            src = src.replace(old, new)
        mod = types.ModuleType(f"{pkgname}.{name}")
        mod.__file__ = path
        mod.__package__ = pkgname
        sys.modules[f"{pkgname}.{name}"] = mod
        exec(compile(src, path, "exec"), mod.__dict__)
        setattr(pkg, name, mod)
        return mod

    try:
        cfg = load("config")
        load("validation")
        core = load("core", patches=[(
            "bp.values[range(len(terminal_states)), range(len(terminal_states))] = 1",
            "bp = pd.DataFrame(np.eye(len(terminal_states)), index=terminal_states, columns=terminal_states)")])
        load("presults")
        utils = load("utils", patches=([("tol=1e-4", f"tol={eig_tol:g}")] if eig_tol else []))
        cfg.JOBLIB_BACKEND = "threading"
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return core, utils
