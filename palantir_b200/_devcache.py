"""Device copies of host arrays this package has just produced.

The reference's API hands NumPy / pandas objects from call to call
(``run_diffusion_maps`` -> ``determine_multiscale_space`` -> ``run_palantir``), so a
drop-in implementation would download every intermediate and upload it again a moment
later.  When a result array leaves the package its device original is remembered here,
keyed by the identity of the host array; the next call that receives the same array
uses the device copy instead of uploading.

This is OPT-IN (``config.REUSE_DEVICE_RESULTS``, default False): host arrays are the
caller's to edit, and whether one was edited in place can only be known for certain by
reading all of it again -- which costs as much as uploading it (48-80 MB at 1 M cells:
6-10 ms either way).  With the switch on, a hit additionally requires a fingerprint of
the host memory (address, shape, dtype, a strided sample of the values) to match, which
catches whole-array edits (a sign flip, a rescaling) but not an edit of a few entries:
only switch it on when the arrays are passed from call to call untouched.  Entries die
with their host array.
"""
from __future__ import annotations

import weakref
from typing import Optional

import numpy as np
import torch

_ENTRIES: dict = {}


def _fingerprint(a: np.ndarray):
    flat = a.reshape(-1) if a.flags.c_contiguous or a.flags.f_contiguous else None
    if flat is None:
        return None
    step = max(1, flat.size // 4096) | 1
    return (a.__array_interface__["data"][0], a.shape, a.strides, a.dtype.str,
            float(np.sum(flat[::step], dtype=np.float64)), float(np.sum(flat[:512], dtype=np.float64)),
            float(np.sum(flat[-512:], dtype=np.float64)))


def remember(host: np.ndarray, dev: torch.Tensor) -> None:
    """``dev`` holds the same numbers as ``host`` (row-major, same shape)."""
    from . import config

    if not config.REUSE_DEVICE_RESULTS or not isinstance(host, np.ndarray) or host.size == 0:
        return
    fp = _fingerprint(host)
    if fp is None:
        return
    key = id(host)
    try:
        ref = weakref.ref(host, lambda _r, k=key: _ENTRIES.pop(k, None))
    except TypeError:
        return
    _ENTRIES[key] = (ref, fp, dev)


def lookup(host) -> Optional[torch.Tensor]:
    """Device copy of ``host`` if one is remembered and the host memory is unchanged."""
    from . import config

    if not config.REUSE_DEVICE_RESULTS or not isinstance(host, np.ndarray):
        return None
    hit = _ENTRIES.get(id(host))
    if hit is None:
        # a view of a remembered array with the same memory layout (DataFrame.values returns a new view object)
        ptr = host.__array_interface__["data"][0]
        for ref, fp, dev in list(_ENTRIES.values()):
            if fp[0] == ptr and fp[1] == host.shape and fp[2] == host.strides and fp[3] == host.dtype.str:
                hit = (ref, fp, dev)
                break
        if hit is None:
            return None
    ref, fp, dev = hit
    if ref() is None:
        return None
    return dev if _fingerprint(host) == fp else None


def clear() -> None:
    _ENTRIES.clear()
