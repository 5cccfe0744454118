"""End-to-end GPU parity of the fused training step through the drop-in module interface.

1. Against the REFERENCE ITSELF: tests/golden/*.npz hold elbo / log_p / kl and every parameter gradient produced by
   running the unmodified reference (oracle/make_golden.py); the CUDA path is run on the same weights, batch and
   noise.
2. Against the fp64 oracle at a larger, MNIST(U)-shaped size, plus size-independent properties at the full cfg1 size
   (data-parallel shard additivity, determinism of the forward ELBO).

Tolerances: ELBO terms are compared at 2e-3 relative (FP16 operands = TF32's 11-bit significand, fp32 accumulation;
measured ~1e-4).  Gradients
are compared at 6e-2 relative Frobenius norm per parameter: a ReLU network's gradient is discontinuous in its
pre-activations, so the TF32-level forward perturbation (the same one the reference's own cuDNN-TF32 path has)
flips ~5e-4 of the activation derivatives and moves gradients by ~sqrt(5e-4) ~ 2e-2; tests/test_gpu_stages.py
shows the backward kernels themselves agree to 5e-3 once the activation pattern is held fixed.
"""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import load_golden, oracle_step, rel_err
from tvae_b200 import synth
from tvae_b200.config import CFG1, CFG2, CFG3, CFG4, HotPathConfig

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = ["g1_mnist", "g2_dsprites", "g3_galaxy", "g4_particles_ctf", "g5_particles_mask", "g6_mnist_noref",
          "g7_particles_fitnoise", "g8_mnist_attn_unimodal", "g9_mnist_resid",
          "g10_mnist_tanh", "g11_particles_tanh", "g12_mnist_attn_unimodal_p4"]


def act_cls(cfg):
    return nn.Tanh if cfg.activation == "tanh" else nn.LeakyReLU


def build_models(cfg, seed=0, gain=1.0):
    import src.models as models
    with contextlib.redirect_stdout(io.StringIO()):
        gen = models.SpatialGenerator(cfg.z, cfg.hidden, n_out=cfg.n_out, num_layers=cfg.gen_layers, activation=act_cls(cfg),
                                      resid=cfg.gen_resid, fourier_expansion=cfg.fourier, sigma=cfg.sigma)
        if cfg.encoder == "attn_unimodal":
            enc = models.InferenceNetwork_AttentionTranslation_UnimodalRotation(cfg.n, cfg.C, cfg.z, kernels_num=cfg.O,
                                                                                activation=act_cls(cfg),
                                                                                groupconv=cfg.G if cfg.G > 1 else 0)
        else:
            enc = models.InferenceNetwork_AttentionTranslation_AttentionRotation(
                cfg.n, cfg.C, cfg.z, kernels_num=cfg.O, kernels_size=cfg.k, padding=cfg.p, activation=act_cls(cfg),
                groupconv=cfg.G, rot_refinement=cfg.rot_refinement, theta_prior=cfg.theta_prior,
                normal_prior_over_r=cfg.normal_prior_over_r)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state(cfg, seed).items()})
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.encoder_state(cfg, seed, gain).items()})
    return gen.to(DEV), enc.to(DEV)


def r_inf_of(cfg):
    if cfg.encoder == "attn_unimodal":
        return "unimodal"
    return "attention+offsets" if cfg.rot_refinement else "attention"


def run_step(cfg, B, seed=0, gain=1.0, backward=True):
    from tvae_b200 import elbo as E
    gen, enc = build_models(cfg, seed, gain)
    data = synth.minibatch(cfg, B, seed)
    nz = {k: torch.from_numpy(v).to(DEV) for k, v in synth.noise(cfg, B, seed).items()}
    x = torch.from_numpy(synth.image_coords(cfg.n)).to(DEV)
    y = torch.from_numpy(data["y"]).to(DEV)
    r_inf = r_inf_of(cfg)
    if cfg.likelihood == "gaussian":
        ctf = torch.from_numpy(data["ctf"]).to(DEV) if data["ctf"] is not None else None
        out = E.eval_minibatch_particles(x, y, ctf, gen, enc, "attention", r_inf, 0, DEV, cfg.theta_prior, cfg.G, cfg.p,
                                         cfg.mask_radius, noise=nz)
    else:
        out = E.eval_minibatch(x, y, gen, enc, "attention", r_inf, 0, DEV, cfg.theta_prior, cfg.G, cfg.n, noise=nz)
    elbo, logp, kl = out
    grads = {}
    if backward:
        (-elbo).backward()
        torch.cuda.synchronize()
        for k, p in enc.named_parameters():
            grads["enc." + k] = p.grad.detach().cpu()
        for k, p in gen.named_parameters():
            grads["gen." + k] = p.grad.detach().cpu()
    return float(elbo.detach()), float(logp.detach()), float(kl.detach()), grads


def check_grads(grads, ref, tol, label):
    worst = ("", 0.0)
    for k, v in grads.items():
        r = torch.as_tensor(ref[k])
        assert tuple(v.shape) == tuple(r.shape), k
        if k == "enc.conv_a.bias":   # exactly zero in exact arithmetic (softmax shift invariance)
            assert float(v.abs().max()) < 1e-3
            continue
        e = rel_err(v, r)
        if e > worst[1]:
            worst = (k, e)
        assert e < tol, (label, k, e)
    print(f"{label}: worst gradient rel err {worst[1]:.2e} ({worst[0]})")


@pytest.mark.parametrize("name", GOLDEN)
def test_step_matches_reference_golden(name):
    g, cfg, B, _ = load_golden(name)
    elbo, logp, kl, grads = run_step(cfg, B)
    print(f"{name}: elbo {elbo:.5f} (ref {g['elbo']:.5f}) log_p {logp:.5f} ({g['log_p']:.5f}) kl {kl:.5f} ({g['kl']:.5f})")
    assert abs(elbo - g["elbo"]) < 2e-3 * abs(g["elbo"])
    assert abs(logp - g["log_p"]) < 2e-3 * abs(g["log_p"])
    assert abs(kl - g["kl"]) < 2e-3 * abs(g["kl"])
    check_grads(grads, {k[5:]: v for k, v in g.items() if k.startswith("grad.")}, 6e-2, name)


def test_step_matches_oracle_mnist_shaped():
    cfg = CFG1.with_(name="cfg1_b8")
    B = 8
    elbo, logp, kl, grads = run_step(cfg, B)
    o_elbo, o_logp, o_kl, _, o_grads = oracle_step(cfg, B, dtype=torch.float64)
    print(f"cfg1 B=8: elbo {elbo:.4f} / {float(o_elbo):.4f}, log_p {logp:.4f} / {float(o_logp):.4f}, kl {kl:.4f} / {float(o_kl):.4f}")
    assert abs(elbo - float(o_elbo)) < 2e-3 * abs(float(o_elbo))
    assert abs(logp - float(o_logp)) < 2e-3 * abs(float(o_logp))
    assert abs(kl - float(o_kl)) < 2e-3 * abs(float(o_kl))
    check_grads(grads, o_grads, 6e-2, "cfg1_b8")


@pytest.mark.parametrize("cfg", [CFG2, CFG3, CFG4], ids=lambda c: c.name)
def test_full_size_step_matches_oracle(cfg):
    """BASELINE.json configs[1..3] at their full image / filter sizes (dSprites 64x64 k=64, galaxy RGB: the one-channel-
    at-a-time slab of the conv kernel, particle stack 128x128 + CTF), B = 2, against the fp32 CPU oracle."""
    B = 2
    elbo, logp, kl, grads = run_step(cfg, B)
    o_elbo, o_logp, o_kl, _, o_grads = oracle_step(cfg, B, dtype=torch.float32)
    print(f"{cfg.name} B=2: elbo {elbo:.4f} / {float(o_elbo):.4f}, log_p {logp:.4f} / {float(o_logp):.4f}, kl {kl:.4f} / {float(o_kl):.4f}")
    assert abs(elbo - float(o_elbo)) < 2e-3 * abs(float(o_elbo))
    assert abs(logp - float(o_logp)) < 2e-3 * abs(float(o_logp))
    assert abs(kl - float(o_kl)) < 2e-3 * abs(float(o_kl))
    check_grads(grads, o_grads, 6e-2, cfg.name)


def test_full_size_properties_cfg1():
    """BASELINE-size batch (B=100): shard additivity (the data-parallel identity) and determinism."""
    cfg, B = CFG1, 100
    e_full, l_full, k_full, g_full = run_step(cfg, B)
    e2, l2, k2, _ = run_step(cfg, B, backward=False)
    assert (e_full, l_full, k_full) == (e2, l2, k2)          # forward ELBO is bit-deterministic
    # mean over two half-batches == full batch (per-image losses, no cross-image coupling)
    from tvae_b200 import elbo as E
    gen, enc = build_models(cfg)
    data = synth.minibatch(cfg, B, 0)
    nz = {k: torch.from_numpy(v).to(DEV) for k, v in synth.noise(cfg, B, 0).items()}
    x = torch.from_numpy(synth.image_coords(cfg.n)).to(DEV)
    y = torch.from_numpy(data["y"]).to(DEV)
    acc = 0.0
    for sl in (slice(0, 50), slice(50, 100)):
        sub = {k: v[sl].contiguous() for k, v in nz.items()}
        e, _, _ = E.eval_minibatch(x, y[sl].contiguous(), gen, enc, "attention", "attention+offsets", 0, DEV, cfg.theta_prior,
                                   cfg.G, cfg.n, noise=sub)
        (-e * 0.5).backward()
        acc += 0.5 * float(e)
    assert abs(acc - e_full) < 1e-5 * abs(e_full)
    torch.cuda.synchronize()
    for k, p in list(enc.named_parameters()) + list(gen.named_parameters()):
        key = ("enc." if k.startswith("conv") else "gen.") + k
        if key == "enc.conv_a.bias":
            continue
        assert rel_err(p.grad.cpu(), g_full[key]) < 2e-3, key    # atomics reorder fp32 sums only


def test_module_interface_matches_oracle():
    """Drop-in module calls (GroupConv / encoder 7-tuple / SpatialGenerator) with torch autograd around them."""
    import src.models as models
    from oracle import target_vae_oracle as orc
    from helpers import oracle_inputs
    cfg = HotPathConfig("mod", C=1, n=24, k=9, p=3, G=8, z=2, O=32, hidden=64)
    B = 2
    gen, enc = build_models(cfg)
    oenc, ogen, x, y, ctf, nz = oracle_inputs(cfg, B, dtype=torch.float64, requires_grad=False)
    yd = y.float().to(DEV)
    out = enc.conv1(yd, DEV)
    ref = orc.groupconv_forward(y, oenc.conv1_w, oenc.conv1_b, cfg.G, cfg.p)
    assert tuple(out.shape) == tuple(ref.shape)
    assert rel_err(out.cpu(), ref) < 3e-3
    bank = enc.conv1.trans_filter(DEV)
    assert rel_err(bank.cpu(), orc.rotated_filter_bank(oenc.conv1_w, cfg.G)) < 6e-4
    torch.manual_seed(0)
    attn, q, p_r, a_s, offs, theta, z = enc(yd, DEV)
    o = orc.encoder_forward(y, oenc, cfg.G, cfg.p, cfg.rot_refinement, cfg.normal_prior_over_r, cfg.theta_prior,
                            torch.zeros(B, cfg.L, dtype=torch.float64))
    for mine, theirs, nm in ((attn, o[0], "attn"), (q, o[1], "q_t_r"), (p_r, o[2], "p_r"), (offs, o[4], "offsets"),
                             (theta, o[5], "theta"), (z, o[6], "z")):
        assert tuple(mine.shape) == tuple(theirs.shape), nm
        assert rel_err(mine.detach().cpu(), theirs) < 3e-3, nm
    assert tuple(a_s.shape) == tuple(o[3].shape) and abs(float(a_s.sum()) - B) < 1e-3
    # generator with autograd to coordinates and z
    xx = x.float().to(DEV).expand(B, -1, -1).contiguous().requires_grad_(True)
    zz = nz["r_z"][:, :, 0].float().to(DEV).requires_grad_(True)
    yh = gen(xx, zz)
    yo = orc.generator_forward(x.expand(B, -1, -1), nz["r_z"][:, :, 0], ogen)
    assert rel_err(yh.detach().cpu(), yo) < 5e-3
    yh.sum().backward()
    assert xx.grad is not None and zz.grad is not None and gen.coord_linear.weight.grad is not None
    # loud failure without CUDA tensors
    with pytest.raises(RuntimeError):
        enc.conv1(y.float(), "cpu")


@pytest.mark.parametrize("B,cfg", [
    (1, HotPathConfig("edge_b1", C=1, n=24, k=9, p=3, G=8, z=2, O=32, hidden=128)),                      # single image
    (3, HotPathConfig("edge_odd", C=1, n=25, k=8, p=2, G=4, z=3, O=64, hidden=256)),                     # odd n, even H', ragged tiles
    (5, HotPathConfig("edge_rgb", C=3, n=16, k=8, p=4, G=8, z=2, O=32, hidden=64, gen_layers=3, n_out=3,
                      likelihood="bernoulli_rgb")),                                                       # RGB, chunk-aligned channels
    (2, HotPathConfig("edge_nofourier", C=1, n=20, k=7, p=3, G=16, z=8, O=128, hidden=512, fourier=False,
                      normal_prior_over_r=True)),                                                         # 19 heads, no Fourier, N % 64 != 0
])
def test_edge_shapes_match_oracle(B, cfg):
    """Ragged / minimal batches and awkward geometries (the reference has no tests; these are the shapes its trainers
    can produce: last minibatch of an epoch, odd crops, RGB input, --groupconv 16 -z 8)."""
    elbo, logp, kl, grads = run_step(cfg, B)
    o_elbo, o_logp, o_kl, _, o_grads = oracle_step(cfg, B, dtype=torch.float64)
    assert abs(elbo - float(o_elbo)) < 2e-3 * abs(float(o_elbo))
    assert abs(logp - float(o_logp)) < 2e-3 * abs(float(o_logp))
    assert abs(kl - float(o_kl)) < 2e-3 * abs(float(o_kl))
    check_grads(grads, o_grads, 6e-2, cfg.name)


def test_empty_batch_fails_loudly():
    from tvae_b200 import elbo as E
    from tvae_b200._lib import TvaeError
    cfg = HotPathConfig("edge_b0", C=1, n=24, k=9, p=3, G=8, z=2, O=32, hidden=128)
    gen, enc = build_models(cfg)
    x = torch.from_numpy(synth.image_coords(cfg.n)).to(DEV)
    y = torch.zeros(0, 1, cfg.n, cfg.n, device=DEV)
    with pytest.raises((TvaeError, RuntimeError, ValueError)):
        E.eval_minibatch(x, y, gen, enc, "attention", "attention+offsets", 0, DEV, cfg.theta_prior, cfg.G, cfg.n)


def test_attention_unimodal_matches_oracle_mnist_shaped():
    """--t-inf attention --r-inf unimodal --groupconv 0 at MNIST's size (Conv2d(1, 128, 28, padding 14): 29 x 29 map),
    theta prior std 0.5 (not pi, so the trainer's theta_prior argument is seen to reach the KL kernel)."""
    cfg = HotPathConfig("cfg1_au", C=1, n=28, k=28, p=14, G=1, z=2, rot_refinement=False, encoder="attn_unimodal",
                        theta_prior=0.5)
    B = 6
    elbo, logp, kl, grads = run_step(cfg, B)
    o_elbo, o_logp, o_kl, _, o_grads = oracle_step(cfg, B, dtype=torch.float64)
    print(f"cfg1_au B=6: elbo {elbo:.4f} / {float(o_elbo):.4f}, log_p {logp:.4f} / {float(o_logp):.4f}, kl {kl:.4f} / {float(o_kl):.4f}")
    assert abs(elbo - float(o_elbo)) < 2e-3 * abs(float(o_elbo))
    assert abs(logp - float(o_logp)) < 2e-3 * abs(float(o_logp))
    assert abs(kl - float(o_kl)) < 2e-3 * abs(float(o_kl))
    check_grads(grads, o_grads, 6e-2, "cfg1_au")


@pytest.mark.parametrize("name", ["g8_mnist_attn_unimodal", "g12_mnist_attn_unimodal_p4"])
def test_attention_unimodal_module_interface(name):
    """InferenceNetwork_AttentionTranslation_UnimodalRotation.forward, groupconv = 0 and groupconv = 4 (rotation pooling):
    the reference's 4-tuple (models.py:319) against the golden outputs of the unmodified reference."""
    g, cfg, B, _ = load_golden(name)
    _, enc = build_models(cfg)
    y = torch.from_numpy(synth.minibatch(cfg, B, 0)["y"]).to(DEV)
    with torch.no_grad():
        attn, a_s, theta, z = enc(y, DEV)
    d = cfg.Hout
    assert tuple(attn.shape) == (B, 1, d, d) and tuple(a_s.shape) == (B, d, d)
    assert tuple(theta.shape) == (B, 2, d, d) and tuple(z.shape) == (B, 2 * cfg.z, d, d)
    for key, t in (("attn", attn), ("theta", theta), ("z", z)):
        ref = torch.from_numpy(g[key])
        assert float((t.cpu() - ref).abs().max()) < 2e-3 * max(1.0, float(ref.abs().max())), key
    assert abs(float(a_s.sum()) - B) < 1e-3


def test_attention_unimodal_pooled_matches_oracle():
    """--r-inf unimodal --groupconv 8 at a larger size (O = 128: tensor-core heads kernels), tanh activation, against the
    fp64 oracle: the rotation pooling kernels (rot_pool_fwd / rot_pool_bwd) with fc_r weights above 1 in magnitude (the
    power-of-two down-scale of the fp16 gradient operand)."""
    cfg = HotPathConfig("cfg1_aup", C=1, n=24, k=24, p=12, G=8, z=2, rot_refinement=False, encoder="attn_unimodal",
                        hidden=128, theta_prior=0.7, activation="tanh")
    B = 4
    from tvae_b200 import elbo as E
    gen, enc = build_models(cfg)
    with torch.no_grad():
        enc.fc_r.weight.mul_(4.0)                     # |w_r| up to ~1.4
    data = synth.minibatch(cfg, B, 0)
    nz = {k: torch.from_numpy(v).to(DEV) for k, v in synth.noise(cfg, B, 0).items()}
    x = torch.from_numpy(synth.image_coords(cfg.n)).to(DEV)
    y = torch.from_numpy(data["y"]).to(DEV)
    elbo, logp, kl = E.eval_minibatch(x, y, gen, enc, "attention", "unimodal", 0, DEV, cfg.theta_prior, cfg.G, cfg.n, noise=nz)
    (-elbo).backward()
    torch.cuda.synchronize()
    from helpers import oracle_inputs, step_config, gen_param_names
    from oracle import target_vae_oracle as orc
    oenc, ogen, ox, oy, _, onz = oracle_inputs(cfg, B, dtype=torch.float64)
    with torch.no_grad():
        oenc.fc_r_w.mul_(4.0)
    o_elbo, o_logp, o_kl, _ = orc.eval_minibatch(ox, oy, oenc, ogen, step_config(cfg), onz["gumbel"], onz["r_z"], onz["r_theta"])
    (-o_elbo).backward()
    assert abs(float(elbo) - float(o_elbo)) < 2e-3 * abs(float(o_elbo))
    assert abs(float(kl) - float(o_kl)) < 2e-3 * abs(float(o_kl))
    ref = {"enc." + n: t.grad for n, t in zip(orc.EncoderParams.names, oenc.tensors())}
    ref.update({"gen." + n: t.grad for (_, t), n in zip(ogen.named_trainable(), gen_param_names(cfg))})
    grads = {"enc." + k: p.grad.detach().cpu() for k, p in enc.named_parameters()}
    grads.update({"gen." + k: p.grad.detach().cpu() for k, p in gen.named_parameters()})
    check_grads(grads, ref, 6e-2, "cfg1_aup")


def test_resid_generator_module_matches_oracle():
    """SpatialGenerator(resid=True).forward / backward (--generator-resid-layers; ResidLinear = act(Wx + b + x),
    models.py:22-30) through the module interface, against the fp64 oracle on explicit coordinates."""
    from helpers import oracle_inputs, gen_param_names
    from oracle import target_vae_oracle as orc
    cfg = HotPathConfig("cfg1_resid", C=1, n=24, k=9, p=2, G=4, z=3, O=32, hidden=128, gen_layers=3, gen_resid=True)
    B = 5
    gen, _ = build_models(cfg)
    _, ogen, x, _, _, nz = oracle_inputs(cfg, B, dtype=torch.float64)
    xb = (x.expand(B, -1, -1) * torch.linspace(0.6, 1.1, B, dtype=torch.float64).view(B, 1, 1)).contiguous()
    zb = nz["r_z"][:, :, 0]
    ref = orc.generator_forward(xb, zb, ogen)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    (ref * w).sum().backward()
    out = gen(xb.float().to(DEV), zb.float().to(DEV))
    (out * w.float().to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(out.detach().cpu(), ref.detach()) < 2e-3
    got = dict(gen.named_parameters())
    for (_, t), name in zip(ogen.named_trainable(), gen_param_names(cfg)):
        assert rel_err(got[name].grad.cpu(), t.grad) < 3e-2, name


@pytest.mark.parametrize("G", [1, 4])
def test_attention_unimodal_module_backward(G):
    """Autograd through InferenceNetwork_AttentionTranslation_UnimodalRotation.forward (EncoderHeadsFn), groupconv = 0
    and groupconv = 4 (fc_r gradients): a fixed linear functional of (attn, theta, z) against the fp64 oracle.  tanh
    activation: no derivative kinks, so the tolerance is the fp16-operand one."""
    from helpers import oracle_inputs
    from oracle import target_vae_oracle as orc
    cfg = HotPathConfig("cfg1_aum", C=1, n=16, k=16, p=8, G=G, z=2, O=64, hidden=32, rot_refinement=False,
                        encoder="attn_unimodal", activation="tanh")
    B = 3
    _, enc = build_models(cfg)
    y = torch.from_numpy(synth.minibatch(cfg, B, 0)["y"]).to(DEV)
    gen = torch.Generator().manual_seed(11)
    d = cfg.Hout
    wa, wt, wz = (torch.randn(B, c, d, d, generator=gen, dtype=torch.float64) for c in (1, 2, 2 * cfg.z))
    attn, _, theta, z = enc(y, DEV)
    ((attn * wa.float().to(DEV)).sum() + (theta * wt.float().to(DEV)).sum() + (z * wz.float().to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    oenc, _, _, oy, _, _ = oracle_inputs(cfg, B, dtype=torch.float64)
    o_attn, o_theta, o_z = orc.plainconv_head_maps(oy, oenc, cfg.p)
    ((o_attn * wa).sum() + (o_theta.squeeze(2) * wt).sum() + (o_z.squeeze(2) * wz).sum()).backward()
    got = dict(enc.named_parameters())
    for name, t in zip(orc.EncoderParams.names, oenc.tensors()):
        assert tuple(got[name].grad.shape) == tuple(t.grad.shape), name
        assert rel_err(got[name].grad.cpu(), t.grad) < 1e-2, (name, rel_err(got[name].grad.cpu(), t.grad))
