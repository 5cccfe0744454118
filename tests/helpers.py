"""Shared helpers for the test-suite: golden loading and oracle drivers (test infrastructure)."""
import json
import os

import numpy as np
import torch

from tvae_b200 import synth
from tvae_b200.config import HotPathConfig
from oracle import target_vae_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    case = json.loads(bytes(g["case"]).decode())
    cfg = HotPathConfig(**case["cfg"])
    return g, cfg, case["B"], case["trainer"]


def oracle_inputs(cfg, B, seed=0, dtype=torch.float32, gain=1.0, requires_grad=True):
    es = {k: torch.from_numpy(v) for k, v in synth.encoder_state(cfg, seed, gain).items()}
    gs = {k: torch.from_numpy(v) for k, v in synth.generator_state(cfg, seed).items()}
    enc = orc.EncoderParams.from_state_dict(es, dtype, activation=cfg.activation)
    gen = orc.GeneratorParams.from_state_dict(gs, cfg.sigma, activation=cfg.activation, resid=cfg.gen_resid, dtype=dtype)
    if requires_grad:
        for t in enc.tensors():
            t.requires_grad_(True)
        for _, t in gen.named_trainable():
            t.requires_grad_(True)
    data = synth.minibatch(cfg, B, seed)
    nz = synth.noise(cfg, B, seed)
    x = torch.from_numpy(synth.image_coords(cfg.n)).to(dtype)
    y = torch.from_numpy(data["y"]).to(dtype)
    ctf = torch.from_numpy(data["ctf"]).to(dtype) if data["ctf"] is not None else None
    noise = {k: torch.from_numpy(v).to(dtype) for k, v in nz.items()}
    return enc, gen, x, y, ctf, noise


def step_config(cfg):
    return orc.StepConfig(G=cfg.G, padding=cfg.p, rot_refinement=cfg.rot_refinement,
                          normal_prior_over_r=cfg.normal_prior_over_r, theta_prior=cfg.theta_prior,
                          likelihood=cfg.likelihood, mask_radius=cfg.mask_radius, encoder=cfg.encoder)


def oracle_step(cfg, B, seed=0, dtype=torch.float32, gain=1.0, backward=True):
    enc, gen, x, y, ctf, nz = oracle_inputs(cfg, B, seed, dtype, gain)
    elbo, logp, kl, inter = orc.eval_minibatch(x, y, enc, gen, step_config(cfg), nz["gumbel"], nz["r_z"],
                                               nz["r_theta"], ctf)
    grads = {}
    if backward:
        (-elbo).backward()
        for n, t in zip(orc.EncoderParams.names, enc.tensors()):
            grads["enc." + n] = t.grad
        gnames = gen_param_names(cfg)
        for (n_o, t), n_ref in zip(gen.named_trainable(), gnames):
            grads["gen." + n_ref] = t.grad
    return elbo, logp, kl, inter, grads


def gen_param_names(cfg):
    """state_dict names of the trainable generator parameters in oracle `named_trainable` order."""
    names = ["coord_linear.weight", "coord_linear.bias", "latent_linear.weight"]
    idx = 1
    for _ in range(1, cfg.gen_layers):
        pre = f"layers.{idx}.linear" if cfg.gen_resid else f"layers.{idx}"
        names += [pre + ".weight", pre + ".bias"]
        idx += 1 if cfg.gen_resid else 2
    names += [f"layers.{idx}.weight", f"layers.{idx}.bias"]
    return names


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / (b.norm() + 1e-300))


class ForcedActivations:
    """Makes the oracle's LeakyReLUs reproduce given activation values (and hence their derivative masks).

    Gradients of a ReLU-family network are discontinuous in the pre-activations: a TF32-level forward
    perturbation flips the sign of a fraction f ~ 5e-4 of near-zero pre-activations, which alone moves any
    back-propagated gradient by ~sqrt(f) ~ 2e-2 relative (the reference's own TF32 conv path has the same
    property).  To test the backward KERNELS at TF32 tolerance the oracle is therefore evaluated on the same
    activation pattern: value := GPU activation, d(out)/d(pre) := slope mask of that activation.
    """

    def __init__(self, acts):
        self.acts = list(acts)

    def __enter__(self):
        from unittest import mock
        acts = self.acts

        def forced(pre, negative_slope=0.01, inplace=False):
            a = acts.pop(0).to(pre.dtype).reshape(pre.shape)
            mask = torch.where(a > 0, torch.ones_like(a), torch.full_like(a, negative_slope))
            lin = pre * mask
            return lin + (a - lin).detach()

        self._p = mock.patch.object(orc.F, "leaky_relu", forced)
        self._p.start()
        return self

    def __exit__(self, *a):
        self._p.stop()
        assert not self.acts, "oracle consumed fewer activations than supplied"
