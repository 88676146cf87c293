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
