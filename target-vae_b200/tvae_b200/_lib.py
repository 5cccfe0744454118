"""ctypes binding of libtvae_b200.so (C ABI declared in include/tvae_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `csrc/build.py`.  There is no fallback:
if it is missing, importing any compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TVAE_LIB: a development build of the SAME library (csrc/build.py with TVAE_PROBE=1), never a different implementation
LIB_PATH = os.environ.get("TVAE_LIB") or os.path.join(_HERE, "libtvae_b200.so")

_lib = None


class TvaeError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TvaeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the TARGET-VAE hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.tvae_last_error.restype = ctypes.c_char_p
        # development A/B knobs of the library (tvae_test_set_knob), e.g. TVAE_KNOBS="2=1,0=9"; unset = default heuristics
        for kv in filter(None, os.environ.get("TVAE_KNOBS", "").split(",")):
            k, v = kv.split("=")
            _lib.tvae_test_set_knob(int(k), int(v))
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().tvae_last_error().decode(errors="replace")
        raise TvaeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
