"""palantir_b200: the diffusion-map and fate-probability hot path of dpeerlab/Palantir,
rebuilt for NVIDIA B200 (sm_100a) behind the reference's own Python API.

``palantir_b200.utils`` and ``palantir_b200.core`` mirror ``palantir.utils`` /
``palantir.core`` for that path (same signatures, defaults, return types and AnnData
keys); the arithmetic runs in hand-written CUDA kernels reached through the C-ABI
declared in ``include/palantir_b200.h``.  There is no CPU fallback.
"""
from . import config, validation, presults  # noqa: F401
from . import utils, core  # noqa: F401
from .presults import PResults  # noqa: F401

__version__ = "0.1.0"

# The functions of the path, by reference module (what install() binds)
_PATH = {
    "utils": ("run_pca", "compute_kernel", "diffusion_maps_from_kernel", "run_diffusion_maps",
              "determine_multiscale_space", "run_magic_imputation", "early_cell", "fallback_terminal_cell",
              "find_terminal_states", "CellNotFoundException"),
    "core": ("run_palantir", "identify_terminal_states", "_max_min_sampling", "_compute_pseudotime"),
    "presults": ("PResults",),
    "validation": ("normalize_cell_identifiers",),
}


def install(force_alias: bool = False):
    """Make this implementation reachable under the reference's own name ("importable as palantir").

    * The reference package is installed and importable: its functions ON THE PATH are re-bound in place --
      ``palantir.utils.run_diffusion_maps`` & co., ``palantir.core.run_palantir``, ``palantir.run_palantir`` ... --
      so existing scripts and notebooks run the B200 kernels without a changed line; everything off the path
      (plotting, io, gene trends ...) stays the reference's.  Returns the patched ``palantir`` module.
    * It is not importable (or ``force_alias``): ``palantir``, ``palantir.utils``, ``palantir.core``,
      ``palantir.presults``, ``palantir.config`` and ``palantir.validation`` are registered in ``sys.modules`` as
      aliases of this package's modules, so ``import palantir`` and ``from palantir.utils import ...`` resolve here.
    """
    import importlib
    import sys

    here = sys.modules[__name__]
    ref = None
    if not force_alias:
        try:
            ref = importlib.import_module("palantir")
        except Exception:  # noqa: BLE001  (not installed, or its own dependencies are missing)
            ref = None
        if ref is here or getattr(ref, "__pb200_alias__", False):
            return ref
    if ref is None:
        here.__pb200_alias__ = True
        sys.modules["palantir"] = here
        for sub in ("utils", "core", "presults", "config", "validation"):
            sys.modules["palantir." + sub] = getattr(here, sub)
        here.run_palantir = core.run_palantir
        return here
    for mod_name, names in _PATH.items():
        target = getattr(ref, mod_name, None)
        if target is None:
            try:
                target = importlib.import_module("palantir." + mod_name)
            except Exception:  # noqa: BLE001
                continue
        source = getattr(here, mod_name)
        for name in names:
            if hasattr(source, name):
                setattr(target, name, getattr(source, name))
                if hasattr(ref, name):                # names the reference re-exports at package level
                    setattr(ref, name, getattr(source, name))
    ref.__pb200_installed__ = True
    return ref
