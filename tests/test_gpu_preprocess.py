"""GPU: particle-stack input pipeline - tvae_ctf_filter / tvae_crop_normalize against the reference's own
outputs (tests/golden/ctf_golden.npz) and against the fp64 oracle at the full 127 x 127 filter size.
Tolerance: fp64 arithmetic on both sides, fp32 result -> 1e-6 x max|filter| absolute (the oracle's FFT and the kernel's
exact DFT differ by fp64 rounding only); normalisation 2e-6 absolute (the reference accumulates in fp32)."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ctf_golden.npz"))
CTF_CASES = [k for k in G.files if k.startswith("ctf_")]


@pytest.mark.parametrize("key", CTF_CASES)
def test_ctf_filter_matches_reference_golden(key):
    from tvae_b200 import preprocess
    n, m, scale = (int(v) for v in key.split("_")[1:])
    ref = G[key]
    out = preprocess.ctf_filter(G["params"][:ref.shape[0]], n, m, scale=scale)
    assert out.is_cuda and tuple(out.shape) == ref.shape and out.dtype == torch.float32
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=1e-6 * float(np.abs(ref).max()))


def test_ctf_filter_full_size_matches_oracle_and_dataframe_input():
    import pandas as pd
    from tvae_b200 import preprocess
    rng = np.random.default_rng(5)
    B = 12
    params = np.stack([rng.uniform(1.0, 3.0, B), np.full(B, 2.7), np.full(B, 300.0), np.full(B, 2.6), np.full(B, 100.0),
                       np.full(B, 10.0), np.zeros(B), rng.uniform(0, 180, B)], 1)
    ref = po.ctf_filter(params, 127, 127)
    out = preprocess.ctf_filter(pd.DataFrame(params, columns=list(preprocess.CTF_COLUMNS)), 127, 127)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=1e-6 * float(np.abs(ref).max()))
    ref2 = po.ctf_filter(params[:3], 64, 48, 2.0)
    out2 = preprocess.ctf_filter(params[:3], 64, 48, scale=2.0)
    np.testing.assert_allclose(out2.cpu().numpy(), ref2, rtol=0, atol=1e-6 * float(np.abs(ref2).max()))
    assert preprocess.ctf_filter(params[:0], 15, 15).shape == (0, 15, 15)          # empty parameter table


def test_crop_normalize_matches_reference_golden():
    from tvae_b200 import preprocess
    out = preprocess.crop_normalize(G["stack"], crop=24)
    np.testing.assert_allclose(out.cpu().numpy(), G["crop24_norm"], rtol=0, atol=2e-6)
    out = preprocess.crop_normalize(G["stack"])
    np.testing.assert_allclose(out.cpu().numpy(), G["norm"], rtol=0, atol=2e-6)
    raw = preprocess.crop_normalize(G["stack"], crop=24, normalize=False)
    assert np.array_equal(raw.cpu().numpy(), po.crop_normalize(G["stack"], 24, normalize=False))
    big = (np.random.default_rng(1).standard_normal((7, 128, 128)) * 5 + 2).astype(np.float32)
    np.testing.assert_allclose(preprocess.crop_normalize(big).cpu().numpy(), po.crop_normalize(big), rtol=0, atol=2e-6)
