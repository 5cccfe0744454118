"""Pins the oracle against outputs of the reference itself (tests/golden/, made by oracle/make_golden.py).

The reference has no tests or golden vectors of its own, so these fixtures - produced by
importing and running the unmodified reference on seeded inputs - are the parity anchor.
"""
import numpy as np
import pytest
import torch

from helpers import load_golden, oracle_inputs, oracle_step, rel_err, step_config
from oracle import target_vae_oracle as orc

CASES = ["g1_mnist", "g2_dsprites", "g3_galaxy", "g4_particles_ctf", "g5_particles_mask", "g6_mnist_noref",
          "g7_particles_fitnoise", "g9_mnist_resid", "g10_mnist_tanh", "g11_particles_tanh"]


# --t-inf attention --r-inf unimodal --groupconv 0: plain Conv2d encoder, 4-tuple module interface
AU_CASES = ["g8_mnist_attn_unimodal", "g12_mnist_attn_unimodal_p4"]


@pytest.mark.parametrize("name", CASES + AU_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_step_matches_reference(name, dtype):
    g, cfg, B, _ = load_golden(name)
    elbo, logp, kl, inter, grads = oracle_step(cfg, B, dtype=dtype)
    # scalars: reference ran in fp32 (kl/elbo promoted to fp64 by numpy-2 leakage)
    assert abs(float(elbo.detach()) - g["elbo"]) <= 2e-5 * abs(g["elbo"])
    assert abs(float(logp.detach()) - g["log_p"]) <= 2e-5 * abs(g["log_p"])
    assert abs(float(kl.detach()) - g["kl"]) <= 2e-5 * abs(g["kl"])
    for k, v in grads.items():
        ref = g["grad." + k]
        assert tuple(v.shape) == ref.shape, k
        if k == "enc.conv_a.bias":
            # softmax is shift-invariant: this gradient is exactly 0 in exact arithmetic, the
            # reference's value is round-off noise
            assert float(v.abs().max()) < 1e-4 and float(np.abs(ref).max()) < 1e-4
            continue
        assert rel_err(v, ref) < 5e-4, (k, rel_err(v, ref))


@pytest.mark.parametrize("name", CASES)
def test_encoder_tuple_matches_reference(name):
    g, cfg, B, _ = load_golden(name)
    enc, gen, x, y, ctf, nz = oracle_inputs(cfg, B, requires_grad=False)
    out = orc.encoder_forward(y, enc, cfg.G, cfg.p, cfg.rot_refinement, cfg.normal_prior_over_r,
                              cfg.theta_prior, nz["gumbel"])
    for key, t in zip(["attn", "q_t_r", "p_r", "a_sampled", "offsets", "theta", "z"], out):
        ref = torch.from_numpy(g[key])
        assert tuple(t.shape) == tuple(ref.shape), key
        assert torch.allclose(t, ref, rtol=1e-4, atol=2e-5), (key, float((t - ref).abs().max()))


@pytest.mark.parametrize("name", CASES)
def test_bank_and_groupconv_match_reference(name):
    g, cfg, B, _ = load_golden(name)
    enc, gen, x, y, ctf, nz = oracle_inputs(cfg, B, requires_grad=False)
    bank = orc.rotated_filter_bank(enc.conv1_w, cfg.G)
    assert float((bank - torch.from_numpy(g["bank"])).abs().max()) < 1e-6
    out = orc.groupconv_forward(y, enc.conv1_w, enc.conv1_b, cfg.G, cfg.p)
    assert torch.allclose(out, torch.from_numpy(g["conv1_out"]), rtol=1e-4, atol=1e-5)
    # slot G/4 of the bank is an exact 90-degree rotation (SURVEY §8 a-1 probe)
    q = cfg.G // 4
    w = enc.conv1_w[:, :, 0]
    assert float((bank[:, q, :, 0] - torch.rot90(w, -1, dims=(-2, -1))).abs().max()) < 1e-6


@pytest.mark.parametrize("name", CASES)
def test_generator_and_get_latent_match_reference(name):
    g, cfg, B, _ = load_golden(name)
    enc, gen, x, y, ctf, nz = oracle_inputs(cfg, B, requires_grad=False)
    out = orc.generator_forward(x.expand(B, -1, -1), nz["r_z"][:, :, 0], gen)
    ref = torch.from_numpy(g["gen_out"])
    assert torch.allclose(out, ref, rtol=2e-4, atol=2e-4), float((out - ref).abs().max())
    zc, th, dx, ind = orc.get_latent(x, y, enc, cfg.G, cfg.p, cfg.rot_refinement, cfg.normal_prior_over_r,
                                     cfg.theta_prior)
    assert torch.allclose(zc, torch.from_numpy(g["latent_z"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(th, torch.from_numpy(g["latent_theta"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(dx, torch.from_numpy(g["latent_dx"]), rtol=1e-4, atol=1e-5)
    ref_ind = torch.from_numpy(g["attn"]).reshape(B, -1).argmax(1)
    assert torch.equal(ind, ref_ind)


def test_translation_grid_even_odd():
    s = torch.tensor(0.1)
    for d in (4, 5):
        gr = orc.translation_grid(d, s).view(d, d, 2)
        half = d // 2
        for i in range(d):
            for j in range(d):
                assert abs(float(gr[i, j, 0]) - (j - half) * 0.1) < 1e-6
                yy = (half - i) if d % 2 else (half - 1 - i)
                assert abs(float(gr[i, j, 1]) - yy * 0.1) < 1e-6


@pytest.mark.parametrize("name", AU_CASES)
def test_plainconv_encoder_and_get_latent_match_reference(name):
    """InferenceNetwork_AttentionTranslation_UnimodalRotation(groupconv=0): module 4-tuple (models.py:319), conv1
    output and clustering_mnist.get_latent's attention/unimodal branch (:81-120)."""
    g, cfg, B, _ = load_golden(name)
    enc, gen, x, y, ctf, nz = oracle_inputs(cfg, B, requires_grad=False)
    assert enc.conv1_w.dim() == (5 if cfg.G > 1 else 4)
    attn, q_t, p_r, a_s, offs, theta, z = orc.plainconv_encoder_forward(y, enc, cfg.p, nz["gumbel"])
    d = cfg.Hout
    for key, t in (("attn", attn), ("a_sampled", a_s.reshape(B, d, d)), ("theta", theta.squeeze(2)), ("z", z.squeeze(2))):
        ref = torch.from_numpy(g[key])
        assert tuple(t.shape) == tuple(ref.shape), key
        assert torch.allclose(t, ref, rtol=1e-4, atol=2e-5), (key, float((t - ref).abs().max()))
    if cfg.G > 1:
        c1 = orc.groupconv_forward(y, enc.conv1_w, enc.conv1_b, cfg.G, cfg.p)
    else:
        c1 = torch.nn.functional.conv2d(y, enc.conv1_w, enc.conv1_b, padding=cfg.p)
    assert torch.allclose(c1, torch.from_numpy(g["conv1_out"]), rtol=1e-4, atol=1e-5)
    zc, th, dx, ind = orc.get_latent(x, y, enc, 1, cfg.p, False, False, cfg.theta_prior, encoder="attn_unimodal")
    assert torch.allclose(zc, torch.from_numpy(g["latent_z"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(th, torch.from_numpy(g["latent_theta"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(dx, torch.from_numpy(g["latent_dx"]), rtol=1e-4, atol=1e-5)
    assert torch.equal(ind, torch.from_numpy(g["attn"]).reshape(B, -1).argmax(1))
