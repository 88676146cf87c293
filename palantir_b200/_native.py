"""ctypes binding of the C-ABI library (include/palantir_b200.h).

The library takes raw device pointers, sizes and a CUDA stream; PyTorch is used
here only to own device memory and streams.  There is no CPU fallback: calling
any compute entry point without the built library or without a CUDA device
raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpalantir_b200.so")

_P, _I, _L, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double

HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "palantir_b200.h")


def parse_header(path: str = HEADER_PATH):
    """Return {name: (restype_code, arg_codes)} from the prototypes in the C header
    (p = pointer, i = int, l = long long, d = double, v = void)."""
    import re

    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
    text = text.replace('extern "C" {', " ").replace("}", " ")
    out = {}

    def code(t: str) -> str:
        t = t.strip()
        if "*" in t:
            return "p"
        t = re.sub(r"\bconst\b|\bunsigned\b", " ", t).split()
        base = " ".join(t[:-1]) if len(t) > 1 else t[0]       # drop the parameter name
        if base in ("long long",):
            return "l"
        if base in ("int",):
            return "i"
        if base in ("double",):
            return "d"
        if base in ("void", ""):
            return "v"
        raise ValueError(f"unsupported type in header: {t!r}")

    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(pb200_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = [a for a in args.split(",") if a.strip() and a.strip() != "void"]
        out[name] = ("p" if "*" in ret else code(ret + " x"), "".join(code(a) for a in args))
    return out


_CODES = {"p": _P, "i": _I, "l": _L, "d": _D}
_lib = None
_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m palantir_b200._build` "
                "(or __graft_entry__.build()). palantir_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (ret, codes) in parse_header().items():
            fn = getattr(lib, name, None)
            if fn is None:
                raise NativeLibraryError(f"{LIB_PATH} does not export {name} (stale build?)")
            fn.restype = {"p": ctypes.c_char_p, "l": ctypes.c_longlong}.get(ret, ctypes.c_int)
            fn.argtypes = [_CODES[c] for c in codes]
        _lib = lib
    return _lib


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise NativeLibraryError(
            "palantir_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def _arg(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise NativeLibraryError("host tensor passed to a device entry point")
        if not a.is_contiguous():
            raise NativeLibraryError("non-contiguous tensor passed to a device entry point")
        return a.data_ptr()
    return a


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


_bound_device = threading.local()


def call(name: str, *args):
    """Call ``name`` with tensors converted to device pointers; the current torch
    stream is appended as the last argument.  Raises on a non-zero return."""
    lib = load()
    dev = torch.cuda.current_device()
    if getattr(_bound_device, "dev", None) != dev:
        rc = lib.pb200_set_device(dev)
        if rc != 0:
            raise NativeLibraryError(lib.pb200_last_error().decode())
        _bound_device.dev = dev
    fn = getattr(lib, name)
    rc = fn(*[_arg(a) for a in args], stream_ptr())
    if rc != 0:
        raise NativeLibraryError(f"{name} failed ({rc}): {lib.pb200_last_error().decode()}")


def launch_count() -> int:
    """Kernel launches issued by the library so far (0 if it was never loaded)."""
    return int(_lib.pb200_launch_count()) if _lib is not None else 0
