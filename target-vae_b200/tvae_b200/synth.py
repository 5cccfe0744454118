"""Deterministic synthetic weights and minibatches for the hot-path configs (numpy only).

Weights follow the reference initialisers' *distributions* (GroupConv: U(+-1/sqrt(C k^2)),
models.py:161-169; Conv3d / Linear: torch default U(+-1/sqrt(fan_in)); Fourier buffers
randn / U(0, 2pi), models.py:42-43) but are drawn from numpy's PCG64 so that the same arrays can be
regenerated on any box from a seed.  Data shapes follow SURVEY.md §8(d).
"""
from __future__ import annotations

import math

import numpy as np

from .config import HotPathConfig


def _u(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def encoder_state(cfg: HotPathConfig, seed: int = 0, gain: float = 1.0) -> dict:
    """state_dict-shaped numpy arrays for InferenceNetwork_AttentionTranslation_AttentionRotation."""
    rng = np.random.default_rng(1000 + seed)
    O, C, k, z = cfg.O, cfg.C, cfg.k, cfg.z
    b1 = 1.0 / math.sqrt(C * k * k)
    bo = 1.0 / math.sqrt(O)
    sd = {
        "conv1.weight": _u(rng, (O, C, 1, k, k), b1) * gain,
        "conv1.bias": _u(rng, (O,), b1),
        "conv2.weight": _u(rng, (O, O, 1, 1, 1), bo) * gain,
        "conv2.bias": _u(rng, (O,), bo),
        "conv_a.weight": _u(rng, (1, O, 1, 1, 1), bo) * gain,
        "conv_a.bias": _u(rng, (1,), bo),
        "conv_r.weight": _u(rng, (2, O, 1, 1, 1), bo),
        "conv_r.bias": _u(rng, (2,), bo),
        "conv_z.weight": _u(rng, (2 * z, O, 1, 1, 1), bo),
        "conv_z.bias": _u(rng, (2 * z,), bo),
    }
    if cfg.encoder == "attn_unimodal":     # nn.Conv2d modules (models.py:281-287): the same draws without the rotation axis
        keep5 = ("conv1.weight",) if cfg.G > 1 else ()          # --groupconv G > 0: conv1 stays a GroupConv
        sd = {k_: (v.reshape(v.shape[0], v.shape[1], *v.shape[3:]) if v.ndim == 5 and k_ not in keep5 else v)
              for k_, v in sd.items()}
        if cfg.G > 1:                      # fc_r = nn.Linear(G, 1), models.py:284
            bg = 1.0 / math.sqrt(cfg.G)
            sd["fc_r.weight"] = _u(rng, (1, cfg.G), bg)
            sd["fc_r.bias"] = _u(rng, (1,), bg)
    return sd


def generator_state(cfg: HotPathConfig, seed: int = 0) -> dict:
    """state_dict-shaped numpy arrays for SpatialGenerator (non-residual layers)."""
    rng = np.random.default_rng(2000 + seed)
    H, E = cfg.hidden, (cfg.fourier_dim if cfg.fourier else 2)
    sd = {}
    if cfg.fourier:
        sd["embed_latent.weight"] = rng.standard_normal((E, 2)).astype(np.float32)
        sd["embed_latent.bias"] = (rng.uniform(0, 1, size=(E,)) * 2 * np.pi).astype(np.float32)
    sd["coord_linear.weight"] = _u(rng, (H, E), 1 / math.sqrt(E))
    sd["coord_linear.bias"] = _u(rng, (H,), 1 / math.sqrt(E))
    sd["latent_linear.weight"] = _u(rng, (H, cfg.z), 1 / math.sqrt(cfg.z))
    idx = 1
    for _ in range(1, cfg.gen_layers):
        # nn.Sequential indices: Linear, activation pairs - or one ResidLinear module per hidden layer (models.py:83-89)
        pre = f"layers.{idx}.linear" if cfg.gen_resid else f"layers.{idx}"
        sd[pre + ".weight"] = _u(rng, (H, H), 1 / math.sqrt(H))
        sd[pre + ".bias"] = _u(rng, (H,), 1 / math.sqrt(H))
        idx += 1 if cfg.gen_resid else 2
    sd[f"layers.{idx}.weight"] = _u(rng, (cfg.n_out, H), 1 / math.sqrt(H))
    sd[f"layers.{idx}.bias"] = _u(rng, (cfg.n_out,), 1 / math.sqrt(H))
    return sd


def image_coords(n: int) -> np.ndarray:
    """x_coord exactly as train_mnist.py:474-479: (n*n, 2) float32."""
    xg = np.linspace(-1, 1, n)
    yg = np.linspace(1, -1, n)
    x0, x1 = np.meshgrid(xg, yg)
    return np.stack([x0.ravel(), x1.ravel()], 1).astype(np.float32)


def _blob(rng, n, size):
    """A random smooth blob of ~size px, rotated and shifted, zero background, values in [0,1]."""
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float64)
    c = (n - 1) / 2
    th = rng.uniform(0, 2 * np.pi)
    sh = rng.normal(0, n / 10.0, size=2)
    xr = (xx - c - sh[0]) * np.cos(th) + (yy - c - sh[1]) * np.sin(th)
    yr = -(xx - c - sh[0]) * np.sin(th) + (yy - c - sh[1]) * np.cos(th)
    img = np.zeros((n, n))
    for _ in range(3):
        m = rng.uniform(-size / 4, size / 4, size=2)
        sx, sy = rng.uniform(size / 10, size / 4, size=2)
        img += np.exp(-0.5 * (((xr - m[0]) / sx) ** 2 + ((yr - m[1]) / sy) ** 2))
    img = img / img.max()
    img[img < 0.2] = 0.0
    return img


def ctf_kernels(rng, B: int, n: int, apix=2.6, kv=300.0, cs_mm=2.7, amp=0.1, bfactor=100.0) -> np.ndarray:
    """Real-space CTF point-spread kernels (B,1,n-1,n-1) fp32 for synthetic particle stacks.

    Same physical model the reference's host preprocessing uses (src/ctf.py:6-55: phase
    gamma = 2pi(-df lam s^2/2 + cs lam^3 s^4/4), amplitude contrast w, B-factor envelope, kernel =
    -fftshift(ifft2(ctf)).real); written independently for synthetic inputs only.
    """
    m = n - 1
    f = np.fft.fftfreq(m) / apix
    fy, fx = np.meshgrid(f, f, indexing="ij")
    s2 = fx ** 2 + fy ** 2
    v = kv * 1e3
    lam = 12.2639 / np.sqrt(v + 0.97845e-6 * v * v)
    cs = cs_mm * 1e7
    out = np.zeros((B, 1, m, m), dtype=np.float32)
    for i in range(B):
        df = rng.uniform(1.0, 3.0) * 1e4
        gamma = 2 * np.pi * (-0.5 * df * lam * s2 + 0.25 * cs * lam ** 3 * s2 ** 2)
        c = (np.sqrt(1 - amp * amp) * np.sin(gamma) - amp * np.cos(gamma)) * np.exp(-bfactor / 4 * s2)
        out[i, 0] = -np.fft.fftshift(np.fft.ifft2(c)).real
    return out


def minibatch(cfg: HotPathConfig, B: int, seed: int = 0) -> dict:
    """{'y': (B,C,n,n) fp32, 'ctf': (B,1,n-1,n-1) or None} on the host."""
    rng = np.random.default_rng(1234 + seed)
    n, C = cfg.n, cfg.C
    y = np.zeros((B, C, n, n), dtype=np.float32)
    ctf = None
    if cfg.likelihood == "bernoulli":
        for b in range(B):
            img = _blob(rng, n, n * 0.56)
            if cfg.name.startswith("cfg2"):
                img = (img > 0.3).astype(np.float64)          # dSprites is binary
            else:
                img = np.round(img * 255) / 255.0              # uint8/255 like MNIST
            y[b, 0] = img
    elif cfg.likelihood == "bernoulli_rgb":
        nhwc = np.zeros((B, n, n, C), dtype=np.float32)
        for b in range(B):
            for c in range(C):
                nhwc[b, :, :, c] = np.round(np.clip(_blob(rng, n, n * 0.5) + rng.normal(0, 0.02, (n, n)), 0, 1) * 255) / 255
        y = nhwc.reshape(B, C, n, n)                           # raw reinterpretation, train_galaxy.py:445-446
    else:
        ctf = ctf_kernels(rng, B, n) if cfg.ctf else None
        for b in range(B):
            img = _blob(rng, n, n * 0.4) * 0.3 + rng.normal(0, 1.0, (n, n))
            img = (img - img.mean()) / img.std()               # --normalize
            y[b, 0] = img
    return dict(y=y, ctf=ctf)


def noise(cfg: HotPathConfig, B: int, seed: int = 0) -> dict:
    """Supplied noise for parity runs: Gumbel (B, L), r_z (B,z,1), r_theta (B,1,1)."""
    rng = np.random.default_rng(4321 + seed)
    e = rng.exponential(1.0, size=(B, cfg.L)).astype(np.float32)
    e = np.maximum(e, np.float32(1e-30))
    return dict(gumbel=(-np.log(e)).astype(np.float32),
                r_z=rng.standard_normal((B, cfg.z, 1)).astype(np.float32),
                r_theta=rng.standard_normal((B, 1, 1)).astype(np.float32))
