"""Hot-path configurations (flag -> ctor mapping of the reference trainers).

Each preset mirrors the defaults of one reference trainer:
  cfg1  train_mnist.py:401-434      cfg2  train_dsprites.py:408-412
  cfg3  train_galaxy.py:415-420     cfg4/5 train_particles.py:481-525
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace


@dataclass(frozen=True)
class HotPathConfig:
    name: str
    C: int            # --in-channels
    n: int            # --image-dim
    k: int            # --encoder-kernel-size
    p: int            # --encoder-padding
    G: int            # --groupconv
    z: int            # -z
    O: int = 128      # --encoder-kernel-number
    hidden: int = 512  # --generator-hidden-dim
    gen_layers: int = 2  # --generator-num-layers
    fourier: bool = True
    fourier_dim: int = 1024       # models.py:74
    n_out: int = 1
    likelihood: str = "bernoulli"  # bernoulli | bernoulli_rgb | gaussian
    rot_refinement: bool = True    # --r-inf attention+offsets
    normal_prior_over_r: bool = False
    theta_prior: float = math.pi
    ctf: bool = False
    mask_radius: int = 0
    batch: int = 100               # --minibatch-size
    # attn_attn: InferenceNetwork_AttentionTranslation_AttentionRotation (--r-inf attention / attention+offsets);
    # attn_unimodal: InferenceNetwork_AttentionTranslation_UnimodalRotation with --groupconv 0 (--r-inf unimodal):
    # k = n, p = n // 2; G = 1 = --groupconv 0 (plain Conv2d(C, O, n, padding n//2)), G > 1 = --groupconv G (P_G group
    # conv pooled over the rotations by fc_r, models.py:281-285, 301-304)
    encoder: str = "attn_attn"
    activation: str = "leakyrelu"  # --activation leakyrelu | tanh, encoder and generator alike (train_mnist.py:516-519)
    gen_resid: bool = False        # --generator-resid-layers: ResidLinear hidden layers (models.py:22-30, 84-86)

    @property
    def Hout(self) -> int:
        return self.n + 2 * self.p - self.k + 1

    @property
    def sigma(self) -> float:
        # train_mnist.py:511 passes 2/(n-1); train_dsprites.py:492-494 leaves the class default 0.01
        return 0.01 if self.name.startswith("cfg2") else 2.0 / (self.n - 1)

    @property
    def L(self) -> int:
        """cells of the attention map: (r, t) for attention/attention, t only for attention/unimodal"""
        return (1 if self.encoder == "attn_unimodal" else self.G) * self.Hout * self.Hout

    def with_(self, **kw) -> "HotPathConfig":
        return replace(self, **kw)

    # algorithmic work per image, SURVEY.md §8(d) (dense MAC count x2, padding taps included)
    def flops_fwd(self) -> dict:
        pos, px = self.Hout ** 2, self.n ** 2
        E, Hd = (self.fourier_dim if self.fourier else 2), self.hidden
        conv1 = 2 * (self.O * self.G) * (self.C * self.k ** 2) * pos
        conv2 = 2 * self.O ** 2 * self.G * pos + 2 * self.O * (3 + 2 * self.z) * self.G * pos
        first = (2 * 2 * E + 2 * E * Hd) if self.fourier else 2 * 2 * Hd
        gen = px * (first + (self.gen_layers - 1) * 2 * Hd * Hd + 2 * Hd * self.n_out) + 2 * self.z * Hd
        ctf = 2 * px * (self.n - 1) ** 2 if self.ctf else 0
        return dict(conv1=conv1, conv2_heads=conv2, generator=gen, ctf=ctf)

    def flops_fwd_bwd(self) -> float:
        f = self.flops_fwd()
        return 2 * f["conv1"] + 3 * f["conv2_heads"] + 3 * f["generator"] + 2 * f["ctf"]


CFG1 = HotPathConfig("cfg1_mnistU", C=1, n=50, k=28, p=8, G=8, z=2)
CFG2 = HotPathConfig("cfg2_dsprites", C=1, n=64, k=64, p=32, G=8, z=2, fourier=False,
                     normal_prior_over_r=True)
CFG3 = HotPathConfig("cfg3_galaxy", C=3, n=64, k=64, p=32, G=8, z=2, gen_layers=4, n_out=3,
                     likelihood="bernoulli_rgb")
CFG4 = HotPathConfig("cfg4_particles", C=1, n=128, k=64, p=16, G=8, z=2, likelihood="gaussian", ctf=True)
CFG4B = HotPathConfig("cfg4b_particles64", C=1, n=64, k=64, p=16, G=8, z=2, likelihood="gaussian", ctf=True)
CFG5 = HotPathConfig("cfg5_particlesP16", C=1, n=128, k=64, p=16, G=16, z=8, likelihood="gaussian", ctf=True,
                     batch=256)

PRESETS = {c.name.split("_")[0]: c for c in (CFG1, CFG2, CFG3, CFG4, CFG4B, CFG5)}


def tiny(base: HotPathConfig, **kw) -> HotPathConfig:
    """Spatially reduced variant used by parity tests (same channel widths)."""
    return base.with_(name=base.name + "_tiny", **kw)
