"""Test infrastructure: drives the UNMODIFIED reference (the plain-Python files `__graft_entry__.build()` copies into the
git-ignored baseline/_ref, which travels to the GPU box) next to the product in one process.

The reference's `src` is a namespace package and the product ships a drop-in package of the same name, so the reference
modules are imported with baseline/_ref first on sys.path and then taken OUT of sys.modules again: the trainer modules
keep their own references (`train_mnist.models` is the reference's `src.models`), while `import src.models` keeps
resolving to the product.  Nothing here is imported by the product.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
from unittest import mock

import torch
import torch.nn as nn

from tvae_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
TRAINER_OF = {"cfg1": "train_mnist", "cfg2": "train_dsprites", "cfg3": "train_galaxy", "cfg4": "train_particles",
              "cfg4b": "train_particles", "cfg5": "train_particles"}
_cache = {}


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "src", "models.py")) and os.path.exists(os.path.join(REF_DIR, "train_mnist.py"))


def import_reference():
    """-> (reference src.models module, {trainer name: module}); cached."""
    if _cache:
        return _cache["models"], _cache["trainers"]
    assert available(), "baseline/_ref is not installed (run __graft_entry__.build() in the build container)"
    with contextlib.suppress(ImportError):
        importlib.import_module("torchvision")      # before `src` is swapped: its torch.library scan walks sys.modules
    saved_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    saved_path = list(sys.path)
    for k in saved_mods:
        del sys.modules[k]
    pkg = os.path.join(ROOT, "target-vae_b200")
    sys.path[:] = [REF_DIR] + [p for p in sys.path if os.path.abspath(p or ".") != pkg]
    try:
        ref_models = importlib.import_module("src.models")
        ref_src = sys.modules["src"]
        if getattr(ref_src, "__file__", None) is None:       # namespace package: give inspect.getfile() something to return
            ref_src.__file__ = os.path.join(REF_DIR, "src", "__init__.py")
        trainers = {}
        for name in sorted(set(TRAINER_OF.values())):
            with contextlib.redirect_stdout(io.StringIO()):
                trainers[name] = importlib.import_module(name)
        assert os.path.abspath(ref_models.__file__).startswith(REF_DIR), ref_models.__file__
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
    _cache["models"], _cache["trainers"] = ref_models, trainers
    return ref_models, trainers


def build_reference_models(cfg, device, seed=0, gain=1.0):
    ref_models, _ = import_reference()
    act = nn.Tanh if cfg.activation == "tanh" else nn.LeakyReLU
    with contextlib.redirect_stdout(io.StringIO()):
        gen = ref_models.SpatialGenerator(cfg.z, cfg.hidden, n_out=cfg.n_out, num_layers=cfg.gen_layers, activation=act,
                                          resid=cfg.gen_resid, fourier_expansion=cfg.fourier, sigma=cfg.sigma)
        enc = ref_models.InferenceNetwork_AttentionTranslation_AttentionRotation(
            cfg.n, cfg.C, cfg.z, kernels_num=cfg.O, kernels_size=cfg.k, padding=cfg.p, activation=act, groupconv=cfg.G,
            rot_refinement=cfg.rot_refinement, theta_prior=cfg.theta_prior, normal_prior_over_r=cfg.normal_prior_over_r)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state(cfg, seed).items()})
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.encoder_state(cfg, seed, gain).items()})
    return gen.to(device), enc.to(device)


class SuppliedNoise:
    """Patches the reference's two RNG draw sites (F.gumbel_softmax at models.py:387, Normal.sample at
    train_mnist.py:206,230) so they consume supplied tensors - "identical inputs" (SURVEY.md 8c)."""

    def __init__(self, noise):
        self.gumbel = noise["gumbel"]
        self.normals = [noise["r_z"], noise["r_theta"]]

    def __enter__(self):
        import torch.nn.functional as F
        from torch.distributions.normal import Normal
        g, normals = self.gumbel, list(self.normals)

        def gumbel_softmax(logits, tau=1, hard=False, eps=1e-10, dim=-1):
            assert tau == 1 and not hard
            return ((logits + g.to(logits)) / tau).softmax(dim)

        def sample(dist, sample_shape=torch.Size()):
            t = normals.pop(0)
            assert tuple(t.shape) == tuple(sample_shape) + (1,), (t.shape, sample_shape)
            return t.clone().to(dist.loc.device)

        self._p = [mock.patch.object(F, "gumbel_softmax", gumbel_softmax), mock.patch.object(Normal, "sample", sample)]
        for p in self._p:
            p.start()
        return self

    def __exit__(self, *a):
        for p in self._p:
            p.stop()


@contextlib.contextmanager
def math_mode(tf32: bool):
    """tf32=False: true fp32 everywhere (the parity oracle); tf32=True: the reference's DEFAULT GPU math mode
    (cudnn.allow_tf32 = True, matmul.allow_tf32 = False: TF32 convolutions, fp32 linears - SURVEY.md 2a)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def reference_step(cfg, B, device, seed=0, gain=1.0, tf32=False, data=None, noise=None):
    """eval_minibatch + (-elbo).backward() of the unmodified reference on `device` with supplied noise.
    -> (elbo, log_p, kl) python floats, {"enc.<name>" / "gen.<name>": grad tensor}."""
    _, trainers = import_reference()
    tm = trainers[TRAINER_OF[cfg.name.split("_")[0]]]
    gen, enc = build_reference_models(cfg, device, seed, gain)
    data = data or synth.minibatch(cfg, B, seed)
    nz = noise or {k: torch.from_numpy(v) for k, v in synth.noise(cfg, B, seed).items()}
    nz = {k: v.to(device) for k, v in nz.items()}
    x = torch.from_numpy(synth.image_coords(cfg.n)).to(device)
    y = torch.from_numpy(data["y"]).to(device)
    r_inf = "attention+offsets" if cfg.rot_refinement else "attention"
    dev = torch.device(device)
    with math_mode(tf32), SuppliedNoise(nz):
        if tm.__name__ == "train_particles":
            ctf = torch.from_numpy(data["ctf"]).to(device) if data["ctf"] is not None else None
            elbo, logp, kl = tm.eval_minibatch(x, y, ctf, gen, enc, "attention", r_inf, 0, dev, cfg.theta_prior, cfg.G, cfg.p,
                                               cfg.mask_radius)
        else:
            elbo, logp, kl = tm.eval_minibatch(x, y, gen, enc, "attention", r_inf, 0, dev, cfg.theta_prior, cfg.G, cfg.n)
        (-elbo).backward()
    grads = {"enc." + k: p.grad.detach() for k, p in enc.named_parameters()}
    grads.update({"gen." + k: p.grad.detach() for k, p in gen.named_parameters()})
    return (float(elbo), float(logp), float(kl)), grads


def reference_attention(enc_ref, y, tf32=False):
    """attn (B, G*H'*W') of the reference encoder's forward (models.py:354-403; includes + p_r), no noise needed."""
    with torch.no_grad(), math_mode(tf32):
        out = enc_ref(y, y.device)
    return out[0].reshape(y.shape[0], -1)
