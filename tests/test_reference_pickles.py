"""Reference checkpoints are whole-module pickles written by the reference's own classes (src/utils.py:42-46,
train_mnist.py:672-684) and restored by class name without running __init__ (clustering_mnist.py:308).  The drop-in
`src.models` must therefore work from exactly the attributes the REFERENCE constructors set.

tests/golden/ref_pickles.pt holds such pickles (written by oracle/make_pickle_fixture.py with the unmodified reference
classes); here they are restored with the product's `src.models` on the path.  CPU part: every attribute the hot path
reads exists.  GPU part: the restored modules produce bit-identical ELBO terms / get_latent outputs to modules built by
the product's own constructors with the same weights.
"""
import io
import os

import pytest
import torch

from helpers import GOLDEN_DIR, load_golden
from tvae_b200 import synth

CASES = ["g1_mnist", "g8_mnist_attn_unimodal", "g9_mnist_resid", "g12_mnist_attn_unimodal_p4"]


def _restore(name):
    import src.models as models              # the product's drop-in: the pickles resolve `src.models.<Name>` to it
    blob = torch.load(os.path.join(GOLDEN_DIR, "ref_pickles.pt"), weights_only=False)[name]
    gen = torch.load(io.BytesIO(blob["generator"]), weights_only=False)
    enc = torch.load(io.BytesIO(blob["encoder"]), weights_only=False)
    assert type(gen) is models.SpatialGenerator and type(enc).__module__ == "src.models"
    return gen, enc


@pytest.mark.parametrize("name", CASES)
def test_restored_modules_expose_hot_path_state(name):
    _, cfg, _, _ = load_golden(name)
    gen, enc = _restore(name)
    # none of these may rely on attributes only the product's __init__ would create
    assert "_sigma" not in gen.__dict__ and "_resid" not in gen.__dict__
    assert gen._resid == cfg.gen_resid
    assert abs(gen._sigma - cfg.sigma) < 1e-6 * cfg.sigma
    assert enc.kernels_size == cfg.k and enc.padding == cfg.p
    spec = enc.encoder_spec()
    assert spec.z == cfg.z and spec.padding == cfg.p
    assert len(enc.hot_path_params()) == spec.n_params
    assert len(gen.hot_path_params()) == 3 + 2 * cfg.gen_layers
    gs, es = synth.generator_state(cfg), synth.encoder_state(cfg)
    for k, v in gen.state_dict().items():
        assert torch.equal(v, torch.from_numpy(gs[k])), k
    for k, v in enc.state_dict().items():
        assert torch.equal(v, torch.from_numpy(es[k])), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_restored_modules_run_the_hot_path(name):
    from test_gpu_step import build_models, r_inf_of
    from tvae_b200 import elbo as E
    dev = "cuda"
    _, cfg, B, _ = load_golden(name)
    gen_r, enc_r = _restore(name)
    gen_r, enc_r = gen_r.to(dev), enc_r.to(dev)
    gen_o, enc_o = build_models(cfg)
    x = torch.from_numpy(synth.image_coords(cfg.n)).to(dev)
    y = torch.from_numpy(synth.minibatch(cfg, B, 0)["y"]).to(dev)
    nz = {k: torch.from_numpy(v).to(dev) for k, v in synth.noise(cfg, B, 0).items()}
    r_inf = r_inf_of(cfg)
    outs = []
    for gen, enc in ((gen_r, enc_r), (gen_o, enc_o)):
        elbo, logp, kl = E.eval_minibatch(x, y, gen, enc, "attention", r_inf, 0, dev, cfg.theta_prior, cfg.G, cfg.n, noise=nz)
        (-elbo).backward()
        lat = E.get_latent(x, y, enc, "attention", r_inf, dev, cfg.n)
        outs.append((elbo, logp, kl, *lat, enc.conv2.weight.grad))
    for a, b in zip(outs[0][:-1], outs[1][:-1]):
        assert torch.equal(a, b)                      # the forward is bit-deterministic
    ga, gb = outs[0][-1], outs[1][-1]                  # gradients are summed with atomics: same up to summation order
    assert float((ga - gb).norm() / gb.norm()) < 1e-5
