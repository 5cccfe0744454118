"""Particle-stack loading: tvae_b200.mrc against files written and parsed by the UNMODIFIED reference
(tests/golden/mrc/*.mrcs + tests/golden/mrc_golden.npz from oracle/make_golden_mrc.py: src/mrc.py `write` / `parse`,
src/image.py `crop`, the --normalize arithmetic of train_particles.py:592-600).

CPU: the host header parser returns the reference's header fields and the rank shards tile the stack exactly.
GPU: every rank's shard, decoded / cropped / standardised on the device, equals the reference's arrays - bit-exact for the
decode, 2e-6 absolute after the standardisation (fp64 arithmetic rounded to fp32 on both sides).
"""
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR
from tvae_b200 import mrc

G = np.load(os.path.join(GOLDEN_DIR, "mrc_golden.npz"))
CASES = sorted({k.split(".")[0] for k in G.files})


def _path(name):
    return os.path.join(GOLDEN_DIR, "mrc", name + ".mrcs")


@pytest.mark.parametrize("name", CASES)
def test_header_matches_reference_parser(name):
    h = mrc.read_header(_path(name))
    nx, ny, nz, mode, nxt = (int(v) for v in G[name + ".header"])
    assert (h.nx, h.ny, h.nz, h.mode, h.next) == (nx, ny, nz, mode, nxt)
    assert (h.nz, h.ny, h.nx) == G[name + ".array"].shape
    assert os.path.getsize(_path(name)) == 1024 + h.next + G[name + ".array"].nbytes


def test_shards_tile_the_stack():
    for n in (0, 1, 7, 8, 9, 100, 1001):
        for world in (1, 2, 3, 8):
            edges = [mrc.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mrc.shard_range(10, 3, 3)


def test_rejects_short_header():
    with pytest.raises(ValueError):
        mrc.parse_header(b"\x00" * 100)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("world", [1, 3])
def test_device_decode_crop_normalize(name, world):
    arr = G[name + ".array"]
    for crop, normalize, key in ((0, False, None), (0, True, ".norm"), (12, True, ".crop12_norm")):
        parts = [mrc.load_stack(_path(name), r, world, crop=crop, normalize=normalize) for r in range(world)]
        assert all(p.is_cuda and p.dtype.is_floating_point for p in parts)
        got = np.concatenate([p.cpu().numpy() for p in parts])
        if key is None:
            assert np.array_equal(got, arr.astype(np.float32))          # decode alone is exact
        else:
            np.testing.assert_allclose(got, G[name + key], rtol=0, atol=2e-6)
