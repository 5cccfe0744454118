"""Host-side logic of the drop-in interface, no GPU needed: parameter names / shapes of the module classes against the
unmodified reference (the golden fixtures hold one gradient per reference parameter), branch routing of
eval_minibatch / get_latent and the loud failures (no CPU fallback, unsupported combinations)."""
import contextlib
import io

import pytest
import torch
import torch.nn as nn

from helpers import load_golden
from tvae_b200 import elbo as E
from tvae_b200 import ops, synth

ALL_GOLDEN = ["g1_mnist", "g2_dsprites", "g3_galaxy", "g4_particles_ctf", "g5_particles_mask", "g6_mnist_noref",
              "g7_particles_fitnoise", "g8_mnist_attn_unimodal", "g9_mnist_resid", "g10_mnist_tanh", "g11_particles_tanh",
              "g12_mnist_attn_unimodal_p4"]


def build(cfg):
    import src.models as models
    act = nn.Tanh if cfg.activation == "tanh" else nn.LeakyReLU
    with contextlib.redirect_stdout(io.StringIO()):
        gen = models.SpatialGenerator(cfg.z, cfg.hidden, n_out=cfg.n_out, num_layers=cfg.gen_layers, activation=act,
                                      resid=cfg.gen_resid, fourier_expansion=cfg.fourier, sigma=cfg.sigma)
        if cfg.encoder == "attn_unimodal":
            enc = models.InferenceNetwork_AttentionTranslation_UnimodalRotation(cfg.n, cfg.C, cfg.z, kernels_num=cfg.O, activation=act,
                                                                                groupconv=cfg.G if cfg.G > 1 else 0)
        else:
            enc = models.InferenceNetwork_AttentionTranslation_AttentionRotation(
                cfg.n, cfg.C, cfg.z, kernels_num=cfg.O, kernels_size=cfg.k, padding=cfg.p, activation=act, groupconv=cfg.G,
                rot_refinement=cfg.rot_refinement, theta_prior=cfg.theta_prior, normal_prior_over_r=cfg.normal_prior_over_r)
    return gen, enc


@pytest.mark.parametrize("name", ALL_GOLDEN)
def test_parameter_names_and_shapes_match_the_reference(name):
    """named_parameters() of the drop-in modules == the reference's (one 'grad.<module>.<name>' entry per reference
    parameter in the fixture), and the seeded state_dicts load strictly."""
    g, cfg, _, _ = load_golden(name)
    gen, enc = build(cfg)
    for prefix, mod in (("grad.enc.", enc), ("grad.gen.", gen)):
        ref = {k[len(prefix):]: v.shape for k, v in g.items() if k.startswith(prefix)}
        mine = {k: tuple(p.shape) for k, p in mod.named_parameters()}
        assert mine == ref, (name, prefix)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state(cfg).items()}, strict=True)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.encoder_state(cfg).items()}, strict=True)
    assert len(enc.hot_path_params()) == enc.encoder_spec().n_params


def test_branch_routing_and_loud_failures():
    _, cfg, _, _ = load_golden("g1_mnist")
    gen, enc = build(cfg)
    _, cfg_u, _, _ = load_golden("g8_mnist_attn_unimodal")
    _, enc_u = build(cfg_u)
    _, cfg_p, _, _ = load_golden("g12_mnist_attn_unimodal_p4")
    _, enc_p = build(cfg_p)
    x = torch.from_numpy(synth.image_coords(cfg.n))
    y = torch.from_numpy(synth.minibatch(cfg, 2)["y"])
    # specs
    es = E._encoder_spec(enc, "attention", "attention+offsets", 3.0)
    assert (es.G, es.attn_G, es.pool, es.theta_prior_std, es.act) == (cfg.G, cfg.G, False, None, ops.ACT_LEAKYRELU)
    es = E._encoder_spec(enc_u, "attention", "unimodal", 0.25)
    assert (es.G, es.attn_G, es.pool, es.theta_prior_std) == (1, 1, False, 0.25)
    es = E._encoder_spec(enc_p, "attention", "unimodal", 0.25)
    assert (es.G, es.attn_G, es.pool, es.n_params) == (4, 1, True, 12) and es.tables() == ([0.0], [0.0])
    # mismatched encoder / branch
    with pytest.raises(ValueError):
        E._encoder_spec(enc, "attention", "unimodal", 1.0)
    with pytest.raises(ValueError):
        E._encoder_spec(enc_u, "attention", "attention", 1.0)
    with pytest.raises(ValueError):
        E._encoder_spec(enc, "attention", "attention", 1.0)          # encoder was built with rot_refinement
    with pytest.raises(NotImplementedError):
        E._encoder_spec(enc, "unimodal", "unimodal", 1.0)
    # no CPU fallback anywhere on the product path
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.eval_minibatch(x, y, gen, enc, "attention", "attention+offsets", 0, "cpu", cfg.theta_prior, cfg.G, cfg.n)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.get_latent(x, y, enc, "attention", "attention+offsets", "cpu", cfg.n)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(y, "cpu")
    # activations: the trainers offer leakyrelu and tanh only
    assert ops.act_kind(nn.Tanh()) == ops.ACT_TANH and ops.act_kind(nn.LeakyReLU()) == ops.ACT_LEAKYRELU
    with pytest.raises(NotImplementedError):
        ops.act_kind(nn.ReLU())
    with pytest.raises(NotImplementedError):
        ops.act_kind(nn.LeakyReLU(0.2))


def test_fit_noise_combinations_the_reference_cannot_run_are_refused():
    _, cfg, _, _ = load_golden("g7_particles_fitnoise")
    gen, enc = build(cfg)
    x = torch.from_numpy(synth.image_coords(cfg.n))
    y = torch.from_numpy(synth.minibatch(cfg, 2)["y"])
    ctf = torch.zeros(2, 1, cfg.n - 1, cfg.n - 1)
    with pytest.raises(NotImplementedError, match="fit-noise"):
        E.eval_minibatch_particles(x, y, ctf, gen, enc, "attention", "attention+offsets", 0, "cpu", cfg.theta_prior, cfg.G, cfg.p, 0)
    with pytest.raises(NotImplementedError, match="fit-noise"):
        E.eval_minibatch_particles(x, y, None, gen, enc, "attention", "attention+offsets", 0, "cpu", cfg.theta_prior, cfg.G, cfg.p, 4)


def test_graphed_step_host_contract():
    """tvae_b200.graph.GraphedStep (one CUDA graph per minibatch): no CPU fallback, a GradSync with a fused optimiser is refused
    (its Adam step count is a host scalar), train_epoch exposes graph=; the header declares the one-bit mask buffer."""
    import inspect
    import os
    from tvae_b200 import train
    from tvae_b200.graph import GraphedStep
    _, cfg, _, _ = load_golden("g1_mnist")
    gen, enc = build(cfg)
    x = torch.from_numpy(synth.image_coords(cfg.n))
    with pytest.raises(RuntimeError):
        GraphedStep(x, (2, cfg.C, cfg.n, cfg.n), gen, enc, "attention", "attention+offsets", "cpu", cfg.theta_prior, cfg.G, cfg.n)

    class FusedSync:
        optimizer = object()
    with pytest.raises((ValueError, RuntimeError, AssertionError)):
        GraphedStep(x, (2, cfg.C, cfg.n, cfg.n), gen, enc, "attention", "attention+offsets", "cuda", cfg.theta_prior, cfg.G, cfg.n,
                    sync=FusedSync())
    assert "graph" in inspect.signature(train.train_epoch).parameters
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "tvae_b200.h")).read()
    assert "mask_bits" in header
    assert [n for n, _ in ops.GenFwdArgs._fields_][-1] == "mask_bits"
