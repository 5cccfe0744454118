"""MRC / MRCS particle stacks onto the device: the host reads the 1024-byte header the way the
reference's `src/mrc.py:108-140 parse` does, every rank uploads only ITS contiguous shard of the image payload (raw
bytes, as stored), and one kernel decodes (modes 0 int8, 1 int16, 2 float32, 6 uint16), centre-crops
(src/image.py:30-42) and standardises (`--normalize`, train_particles.py:592-600) it - the array transforms
`train_particles.py:534-600` performs on the host before building its TensorDataset.

    header = mrc.read_header(path)
    images = mrc.load_stack(path, rank, world, crop=0, normalize=True, device="cuda")      # (N/world, n, m) fp32 on the device

The complex / RGB modes (3, 4, 16) are not particle stacks and are rejected.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
import struct
from collections import namedtuple

import numpy as np
import torch

from . import ops
from ._lib import check, stream_ptr

# the 256 words of the header, in the reference's field order (src/mrc.py:10-104)
_FIELDS = ("nx ny nz mode nxstart nystart nzstart mx my mz xlen ylen zlen alpha beta gamma mapc mapr maps amin amax amean ispg "
           "next creatid nint nreal imodStamp imodFlags idtype lens nd1 nd2 vd1 vd2 tilt_ox tilt_oy tilt_oz tilt_cx tilt_cy tilt_cz "
           "xorg yorg zorg cmap stamp rms nlabl labels").split()
_STRUCT = struct.Struct("<3ii3i3i3f3f3i3f2ih30x2h20x2i6h6f3f4s4sfi800s")
MRCHeader = namedtuple("MRCHeader", _FIELDS)
_DTYPE = {0: np.int8, 1: np.int16, 2: np.float32, 6: np.uint16}
assert _STRUCT.size == 1024


def parse_header(content: bytes) -> MRCHeader:
    """The first 1024 bytes of an MRC file -> header fields (src/mrc.py:110-111)."""
    if len(content) < 1024:
        raise ValueError("MRC header needs 1024 bytes")
    return MRCHeader._make(_STRUCT.unpack(bytes(content[:1024])))


def read_header(path) -> MRCHeader:
    with open(path, "rb") as f:
        return parse_header(f.read(1024))


def _lib():
    lib = ops.L()
    if not getattr(lib, "_tvae_mrc_configured", False):
        lib.tvae_mrc_crop_normalize.restype = ctypes.c_int
        lib.tvae_mrc_crop_normalize.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib._tvae_mrc_configured = True
    return lib


def shard_range(n_images: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank `rank`: the first n_images % world ranks hold one image more."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def load_stack(path, rank: int = 0, world: int = 1, crop: int = 0, normalize: bool = False, device="cuda"):
    """-> (hi - lo, c, c) fp32 CUDA tensor: this rank's shard of the stack, decoded / cropped / standardised on the device.
    Only the shard's bytes are read from the file (memory map) and copied to the GPU."""
    h = read_header(path)
    if h.mode not in _DTYPE:
        raise NotImplementedError(f"MRC mode {h.mode} is not an image stack (supported: 0 int8, 1 int16, 2 float32, 6 uint16)")
    dt = np.dtype(_DTYPE[h.mode])
    nx, ny, nz = h.nx, h.ny, h.nz
    if min(nx, ny, nz) < 1:
        raise ValueError(f"bad MRC dimensions {(nz, ny, nx)}")
    lo, hi = shard_range(nz, rank, world)
    per = ny * nx * dt.itemsize
    raw = np.memmap(path, dtype=np.uint8, mode="r", offset=1024 + h.next + lo * per, shape=((hi - lo) * per,))
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("load_stack: the device path needs a CUDA device (no CPU fallback)")
    buf = torch.from_numpy(np.ascontiguousarray(raw)).to(dev)
    c0, c1 = (crop, crop) if crop > 0 else (ny, nx)
    out = torch.empty(hi - lo, c0, c1, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib().tvae_mrc_crop_normalize(buf.data_ptr(), h.mode, hi - lo, ny, nx, int(crop), 1 if normalize else 0, out.data_ptr(),
                                             stream_ptr()), "tvae_mrc_crop_normalize")
    return out
