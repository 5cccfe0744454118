"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

For each case the reference modules are constructed with the reference constructors, their
state_dict is overwritten with seeded numpy weights (tvae_b200.synth), and the reference
`train_*.eval_minibatch` + `(-elbo).backward()` / `clustering_mnist.get_latent` is run.  Noise is made
"identical inputs" by patching the two RNG entry points the reference draws from
(`F.gumbel_softmax` at models.py:387 and `Normal.sample` at train_mnist.py:206,230) so they consume
supplied tensors (SURVEY.md §8c "Noise control").  Nothing else of the reference is altered.
Fixtures hold: the case description (from which inputs/weights/noise are regenerated), the
reference's scalar outputs, selected intermediates and every parameter gradient.
"""
from __future__ import annotations

import io
import contextlib
import json
import os
import sys
from unittest import mock

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "target-vae_b200"))
REF = os.environ.get("TVAE_REFERENCE", "/root/reference")

from tvae_b200 import synth                      # noqa: E402
from tvae_b200.config import HotPathConfig       # noqa: E402


def _import_reference():
    for m in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "seaborn",
              "astropy", "astropy.stats", "astropy.units"):
        sys.modules.setdefault(m, mock.MagicMock())
    # the product package also ships a `src` mirror (a regular package, which would shadow the reference's
    # namespace package `src`): take it off sys.path while the reference is imported
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    pkg = os.path.join(ROOT, "target-vae_b200")
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != pkg]
    sys.path.insert(0, REF)
    import src.models as ref_models
    import train_mnist, train_dsprites, train_galaxy, train_particles, clustering_mnist
    assert ref_models.__file__.startswith(REF), ref_models.__file__
    return ref_models, dict(mnist=train_mnist, dsprites=train_dsprites, galaxy=train_galaxy,
                            particles=train_particles), clustering_mnist


CASES = {
    # name: (trainer, cfg, B)
    "g1_mnist": ("mnist", HotPathConfig("cfg1_g", C=1, n=20, k=9, p=3, G=8, z=2, O=32, hidden=64), 3),
    "g2_dsprites": ("dsprites", HotPathConfig("cfg2_g", C=1, n=16, k=16, p=8, G=4, z=2, O=32, hidden=64,
                                              fourier=False, normal_prior_over_r=True), 2),
    "g3_galaxy": ("galaxy", HotPathConfig("cfg3_g", C=3, n=12, k=12, p=6, G=8, z=3, O=32, hidden=32,
                                          gen_layers=4, n_out=3, likelihood="bernoulli_rgb"), 2),
    "g4_particles_ctf": ("particles", HotPathConfig("cfg4_g", C=1, n=16, k=9, p=2, G=16, z=8, O=32, hidden=32,
                                                    likelihood="gaussian", ctf=True), 3),
    "g5_particles_mask": ("particles", HotPathConfig("cfg4_gm", C=1, n=16, k=9, p=2, G=8, z=2, O=32, hidden=32,
                                                     likelihood="gaussian", ctf=True, mask_radius=5), 2),
    "g6_mnist_noref": ("mnist", HotPathConfig("cfg1_gn", C=1, n=14, k=7, p=2, G=4, z=2, O=32, hidden=32,
                                              rot_refinement=False), 2),
    # --t-inf attention --r-inf unimodal --groupconv 0 (train_mnist.py:88-183, models.py:268-319): plain Conv2d encoder
    "g8_mnist_attn_unimodal": ("mnist", HotPathConfig("cfg1_gu", C=1, n=16, k=16, p=8, G=1, z=2, O=32, hidden=64,
                                                      rot_refinement=False, encoder="attn_unimodal"), 3),
    # --generator-resid-layers (train_mnist.py:422,508, models.py:22-30, 84-86): three ResidLinear hidden layers
    "g9_mnist_resid": ("mnist", HotPathConfig("cfg1_gr", C=1, n=14, k=7, p=2, G=4, z=2, O=32, hidden=64, gen_layers=4,
                                              gen_resid=True), 3),
    # --activation tanh (train_mnist.py:423,516-519), encoder and generator
    "g10_mnist_tanh": ("mnist", HotPathConfig("cfg1_gt", C=1, n=14, k=7, p=2, G=4, z=2, O=32, hidden=64, gen_layers=3,
                                              activation="tanh"), 3),
    "g11_particles_tanh": ("particles", HotPathConfig("cfg4_gt", C=1, n=16, k=9, p=2, G=8, z=2, O=32, hidden=32,
                                                      likelihood="gaussian", ctf=True, activation="tanh"), 2),
    # --t-inf attention --r-inf unimodal --groupconv 4: P4 group conv pooled over the rotations by fc_r (models.py:281-285, 301-304)
    "g12_mnist_attn_unimodal_p4": ("mnist", HotPathConfig("cfg1_gup", C=1, n=16, k=16, p=8, G=4, z=2, O=32, hidden=64,
                                                          rot_refinement=False, encoder="attn_unimodal"), 3),
    # --fit-noise (train_particles.py:663-666): generator n_out = 2, learned per-pixel log-variance, no CTF / mask
    "g7_particles_fitnoise": ("particles", HotPathConfig("cfg4_gf", C=1, n=16, k=9, p=2, G=8, z=2, O=32, hidden=32,
                                                         likelihood="gaussian", n_out=2), 3),
}


def act_cls(cfg):
    return nn.Tanh if cfg.activation == "tanh" else nn.LeakyReLU


def build_reference_models(ref_models, cfg: HotPathConfig, seed=0):
    with contextlib.redirect_stdout(io.StringIO()):
        gen = ref_models.SpatialGenerator(cfg.z, cfg.hidden, n_out=cfg.n_out, num_layers=cfg.gen_layers,
                                          activation=act_cls(cfg), resid=cfg.gen_resid,
                                          fourier_expansion=cfg.fourier, sigma=cfg.sigma)
        if cfg.encoder == "attn_unimodal":
            assert cfg.k == cfg.n and cfg.p == cfg.n // 2
            enc = ref_models.InferenceNetwork_AttentionTranslation_UnimodalRotation(
                cfg.n, cfg.C, cfg.z, kernels_num=cfg.O, activation=act_cls(cfg), groupconv=cfg.G if cfg.G > 1 else 0)
        else:
            enc = ref_models.InferenceNetwork_AttentionTranslation_AttentionRotation(
                cfg.n, cfg.C, cfg.z, kernels_num=cfg.O, kernels_size=cfg.k, padding=cfg.p, activation=act_cls(cfg),
                groupconv=cfg.G, rot_refinement=cfg.rot_refinement, theta_prior=cfg.theta_prior,
                normal_prior_over_r=cfg.normal_prior_over_r)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state(cfg, seed).items()})
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.encoder_state(cfg, seed).items()})
    return gen, enc


class SuppliedNoise:
    """Patches the reference's two RNG draw sites to return supplied tensors."""

    def __init__(self, gumbel, r_z, r_theta):
        self.gumbel = torch.from_numpy(gumbel)
        self.normals = [torch.from_numpy(r_z), torch.from_numpy(r_theta)]

    def __enter__(self):
        import torch.nn.functional as F
        from torch.distributions.normal import Normal
        g = self.gumbel
        normals = list(self.normals)

        def gumbel_softmax(logits, tau=1, hard=False, eps=1e-10, dim=-1):
            assert tau == 1 and not hard
            return ((logits + g.to(logits)) / tau).softmax(dim)

        def sample(dist, sample_shape=torch.Size()):
            t = normals.pop(0)
            assert tuple(t.shape) == tuple(sample_shape) + (1,), (t.shape, sample_shape)
            return t.clone()

        self._p = [mock.patch.object(F, "gumbel_softmax", gumbel_softmax),
                   mock.patch.object(Normal, "sample", sample)]
        for p in self._p:
            p.start()
        return self

    def __exit__(self, *a):
        for p in self._p:
            p.stop()


def run_case(name, ref_models, trainers, clustering):
    trainer, cfg, B = CASES[name]
    gen, enc = build_reference_models(ref_models, cfg)
    data = synth.minibatch(cfg, B, seed=0)
    nz = synth.noise(cfg, B, seed=0)
    x = torch.from_numpy(synth.image_coords(cfg.n))
    y = torch.from_numpy(data["y"])
    dev = torch.device("cpu")
    tm = trainers[trainer]
    r_inf = "attention+offsets" if cfg.rot_refinement else "attention"
    unimodal = cfg.encoder == "attn_unimodal"
    if unimodal:
        r_inf = "unimodal"
    with SuppliedNoise(nz["gumbel"], nz["r_z"], nz["r_theta"]):
        if trainer == "particles":
            ctf = torch.from_numpy(data["ctf"]) if data["ctf"] is not None else None
            elbo, logp, kl = tm.eval_minibatch(x, y, ctf, gen, enc, "attention", r_inf, 0, dev,
                                               cfg.theta_prior, cfg.G, cfg.p, cfg.mask_radius)
        else:
            elbo, logp, kl = tm.eval_minibatch(x, y, gen, enc, "attention", r_inf, 0, dev,
                                               cfg.theta_prior, cfg.G, cfg.n)
    (-elbo).backward()
    out = {
        "case": np.frombuffer(json.dumps(dict(trainer=trainer, B=B, cfg=cfg.__dict__)).encode(), dtype=np.uint8),
        "elbo": np.float64(elbo.item()), "log_p": np.float64(logp.item()), "kl": np.float64(kl.item()),
        "elbo_dtype": np.frombuffer(str(elbo.dtype).encode(), dtype=np.uint8),
    }
    for k, p in enc.named_parameters():
        out["grad.enc." + k] = p.grad.numpy().astype(np.float32)
    for k, p in gen.named_parameters():
        out["grad.gen." + k] = p.grad.numpy().astype(np.float32)
    # encoder intermediates (module-interface contract, models.py:403) with the same Gumbel noise
    with torch.no_grad(), SuppliedNoise(nz["gumbel"], nz["r_z"], nz["r_theta"]):
        if unimodal:      # 4-tuple of models.py:319
            attn, a_s, theta, z = enc(y, dev)
            out.update(attn=attn.numpy(), a_sampled=a_s.numpy(), theta=theta.numpy(), z=z.numpy())
        else:
            attn, q, p_r, a_s, offs, theta, z = enc(y, dev)
            out.update(attn=attn.numpy(), q_t_r=q.numpy(), p_r=p_r.numpy(), a_sampled=a_s.numpy(),
                       offsets=offs.numpy(), theta=theta.numpy(), z=z.numpy())
    # rotated filter bank and group conv output (a-1, a-2)
    with torch.no_grad():
        if unimodal:
            out["conv1_out"] = (enc.conv1(y, dev) if cfg.G > 1 else enc.conv1(y)).numpy()
        else:
            out["bank"] = enc.conv1.trans_filter(dev).numpy()
            out["conv1_out"] = enc.conv1(y, dev).numpy()
        zc, th, dx = clustering.get_latent(x, y, enc, "attention", r_inf, dev, cfg.n)
        out.update(latent_z=zc.numpy(), latent_theta=th.numpy(), latent_dx=dx.numpy())
        # generator alone on the untransformed grid (a-5, a-6)
        zb = torch.from_numpy(nz["r_z"][:, :, 0])
        out["gen_out"] = gen(x.expand(B, -1, -1).contiguous(), zb).numpy()
    return out


def main():
    ref_models, trainers, clustering = _import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in (sys.argv[1:] or CASES):     # optional case names: regenerate only those fixtures
        out = run_case(name, ref_models, trainers, clustering)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: elbo={out['elbo']:.6f} log_p={out['log_p']:.6f} kl={out['kl']:.6f} "
              f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
