"""TEST INFRASTRUCTURE ONLY - CPU restatement of the particle-stack input pipeline:

  * ctf_filter      real-space CTF point-spread filters from per-micrograph parameters
                    (src/ctf.py:6-23 `compute_2d_ctf`, src/ctf.py:32-55 `ctf_filter`; called at train_particles.py:543-547)
  * crop_normalize  centre crop (src/image.py:30-42, train_particles.py:584-587) followed by the per-image
                    standardisation of `--normalize` (train_particles.py:592-600; population std, ddof = 0)

Vectorised over images and written independently of the reference's loops; pinned by tests/test_preprocess_oracle.py
against tests/golden/ctf_golden.npz, which holds outputs of the UNMODIFIED reference functions run in the build
container (oracle/make_golden_ctf.py).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

COLUMNS = ("defocus", "cs", "voltage", "apix", "bfactor", "ampcont", "dfdiff", "dfang")   # src/ctf.py:28


def ctf_filter(params, n: int, m: int, scale: float = 1.0) -> np.ndarray:
    """params (B, 8) in COLUMNS order -> (B, n, m) float32 = -fftshift(ifft2(CTF)).real  (src/ctf.py:32-55)."""
    p = np.asarray(params, dtype=np.float64).reshape(-1, 8)
    defocus, cs, volt, apix, bfac, ampcont, _dfdiff, dfang = (p[:, i][:, None, None] for i in range(8))
    fx = np.fft.fftfreq(n)[None, :, None] / (apix * scale)          # freqs[:, 0] / apix   (src/ctf.py:35-44)
    fy = np.fft.fftfreq(m)[None, None, :] / (apix * scale)
    dfu = dfv = defocus * 10000.0                                   # src/ctf.py:45-46
    ang0 = 2.0 * np.pi * dfang / 360.0
    w = ampcont / 100.0
    kv = volt * 1000.0                                              # src/ctf.py:8-9
    csa = cs * 1e7
    lam = 12.2639 / np.sqrt(kv + 0.97845e-6 * kv ** 2)              # src/ctf.py:12
    ang = np.arctan2(fy + 0.0 * fx, fx + 0.0 * fy)
    s2 = fx ** 2 + fy ** 2
    df = 0.5 * (dfu + dfv + (dfu - dfv) * np.cos(2.0 * (ang - ang0)))
    gamma = 2.0 * np.pi * (-0.5 * df * lam * s2 + 0.25 * csa * lam ** 3 * s2 ** 2)
    ctf = np.sqrt(1.0 - w ** 2) * np.sin(gamma) - w * np.cos(gamma)
    ctf = ctf * np.exp(-bfac / 4.0 * s2)                            # src/ctf.py:20-21
    real = np.fft.ifft2(ctf, axes=(1, 2)).real
    return (-np.fft.fftshift(real, axes=(1, 2))).astype(np.float32)  # src/ctf.py:53


def crop_normalize(stack, crop: int = 0, normalize: bool = True) -> np.ndarray:
    """(N, n, m) -> centre crop to (N, crop, crop) when crop > 0, then (x - mean) / std per image (ddof = 0)."""
    x = np.asarray(stack, dtype=np.float64)
    if crop > 0:
        n, m = x.shape[-2:]
        si, sj = (n - crop) // 2, (m - crop) // 2                    # src/image.py:36-40
        x = x[..., si:si + crop, sj:sj + crop]
    if normalize:
        flat = x.reshape(x.shape[0], -1)
        mu = flat.mean(1)[:, None, None]
        sd = flat.std(1)[:, None, None]                             # numpy default ddof = 0 (train_particles.py:594-595)
        x = (x - mu) / sd
    return x.astype(np.float32)
