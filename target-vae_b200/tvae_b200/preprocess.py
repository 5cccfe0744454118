"""Particle-stack input pipeline on the device: the two array transforms the particle trainer
performs between loading the `.mrcs` stack and building its `TensorDataset` (train_particles.py:534-600).

  ctf_filter(ctf_params, n, m, scale=1)    src/ctf.py:32-55 - same arguments (the DataFrame `parse_ctf` returns, or an
                                            (N, 8) array in its column order); returns a CUDA tensor (N, n, m) fp32
  crop_normalize(stack, crop=0, normalize=True)   src/image.py:30-42 + train_particles.py:592-600

Parsing the MRC container and the CTF parameter table stays host code (`src/mrc.py`, `src/ctf.py:26-29` are file
I/O).  No CPU fallback: inputs are moved to the CUDA device.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import ops
from ._lib import check, stream_ptr

CTF_COLUMNS = ("defocus", "cs", "voltage", "apix", "bfactor", "ampcont", "dfdiff", "dfang")


def _lib():
    lib = ops.L()
    if not getattr(lib, "_tvae_pre_configured", False):
        lib.tvae_ctf_filter.restype = ctypes.c_int
        lib.tvae_ctf_filter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                        ctypes.c_void_p, ctypes.c_void_p]
        lib.tvae_crop_normalize.restype = ctypes.c_int
        lib.tvae_crop_normalize.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib._tvae_pre_configured = True
    return lib


def _params_array(ctf_params) -> np.ndarray:
    if hasattr(ctf_params, "columns"):                      # pandas DataFrame from src.ctf.parse_ctf
        return np.stack([np.asarray(ctf_params[c], dtype=np.float64) for c in CTF_COLUMNS], 1)
    a = np.asarray(ctf_params, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 8:
        raise ValueError("ctf_params must be parse_ctf's DataFrame or an (N, 8) array in its column order")
    return a


def ctf_filter(ctf_params, n, m, scale=1, device="cuda"):
    p = torch.from_numpy(np.ascontiguousarray(_params_array(ctf_params))).to(device)
    if not p.is_cuda:
        raise RuntimeError("ctf_filter: the device path needs a CUDA device (no CPU fallback)")
    out = torch.empty(p.shape[0], int(n), int(m), device=p.device, dtype=torch.float32)
    with torch.cuda.device(p.device):
        check(_lib().tvae_ctf_filter(p.data_ptr(), p.shape[0], int(n), int(m), float(scale), out.data_ptr(), stream_ptr()),
              "tvae_ctf_filter")
    return out


def crop_normalize(stack, crop=0, normalize=True, device="cuda"):
    x = torch.as_tensor(stack)
    x = x.to(device=device, dtype=torch.float32).contiguous()
    if not x.is_cuda:
        raise RuntimeError("crop_normalize: the device path needs a CUDA device (no CPU fallback)")
    if x.dim() != 3:
        raise ValueError("crop_normalize expects a stack (N, n, m)")
    N, n, m = x.shape
    c0, c1 = (crop, crop) if crop > 0 else (n, m)
    out = torch.empty(N, c0, c1, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib().tvae_crop_normalize(x.data_ptr(), N, n, m, int(crop), 1 if normalize else 0, out.data_ptr(), stream_ptr()),
              "tvae_crop_normalize")
    return out
