"""Generates tests/golden/mrc/*.mrcs and tests/golden/mrc_golden.npz with the UNMODIFIED reference (src/mrc.py `make_header`
/ `write` / `parse`, src/image.py `crop`, the --normalize arithmetic of train_particles.py:592-600) in the build container.

    python oracle/make_golden_mrc.py        # needs /root/reference; the fixtures are committed, the reference is not

One small stack per supported MRC mode (0 int8, 1 int16, 2 float32, 6 uint16), one of them with an extended header; for
each: the file the reference's writer produced, the array and header fields its parser returns, and the cropped +
standardised stack the particle trainer would build from it.
"""
import os
import sys

import numpy as np

REF = os.environ.get("TVAE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import src.image as I        # noqa: E402
import src.mrc as M          # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "tests", "golden", "mrc")
os.makedirs(DST, exist_ok=True)
rng = np.random.default_rng(2024)
CASES = {   # name: (dtype, shape (nz, ny, nx), extended header bytes)
    "mode0_int8": (np.int8, (5, 20, 18), 0),
    "mode1_int16": (np.int16, (7, 24, 24), 80),
    "mode2_float32": (np.float32, (6, 32, 28), 0),
    "mode6_uint16": (np.uint16, (9, 16, 16), 0),
}
out = {}
for name, (dt, shape, ext) in CASES.items():
    if dt == np.float32:
        arr = (rng.standard_normal(shape) * 3.0 + 1.5).astype(dt)
    else:
        info = np.iinfo(dt)
        arr = rng.integers(max(info.min, -3000), min(info.max, 3000), size=shape, endpoint=True).astype(dt)
    header = M.make_header(arr.shape, (1, 1, 1), (90, 90, 90), dtype=dt, exthd_size=ext)
    path = os.path.join(DST, name + ".mrcs")
    with open(path, "wb") as f:
        M.write(f, arr, header=header, extended_header=bytes(range(ext)))
    with open(path, "rb") as f:
        parsed, hdr, exthd = M.parse(f.read())
    assert np.array_equal(parsed, arr) and len(exthd) == ext
    out[name + ".array"] = parsed
    out[name + ".header"] = np.array([hdr.nx, hdr.ny, hdr.nz, hdr.mode, hdr.next], dtype=np.int64)
    crop = 12
    c = I.crop(parsed, crop)
    mu = c.reshape(-1, crop * crop).mean(1)
    std = c.reshape(-1, crop * crop).std(1)
    out[name + ".crop12_norm"] = ((c - mu[:, np.newaxis, np.newaxis]) / std[:, np.newaxis, np.newaxis]).astype(np.float32)
    n, m = parsed.shape[1:]
    mu = parsed.reshape(-1, n * m).mean(1)
    std = parsed.reshape(-1, n * m).std(1)
    out[name + ".norm"] = ((parsed - mu[:, np.newaxis, np.newaxis]) / std[:, np.newaxis, np.newaxis]).astype(np.float32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mrc_golden.npz"), **out)
print("wrote", DST, {k: v.shape for k, v in out.items()})
