"""Per-stage GPU parity: each C-ABI call against the CPU oracle (fp64) on the same seeded inputs.

Tolerances (stated per stage):
  * tensor-core stages (group conv, 1x1x1 convs, generator linears): FP16 operands (11-bit significand, 2^-11 relative - the
    precision class of the reference's cuDNN-TF32 path),
    fp32 accumulate -> relative Frobenius error <= 3e-3 on activations and <= 5e-3 on gradients when the
    oracle is evaluated on the same LeakyReLU activation pattern (helpers.ForcedActivations explains why the
    un-forced comparison is kink-limited to ~3e-2; that end-to-end bound is asserted in test_gpu_step.py).
  * fp32 CUDA-core stages (filter bank, attention/KL, likelihoods, coordinate transforms): <= 1e-4 relative
    (fast-math exp/log differences only), filter bank additionally carries the fp16 rounding of its output.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import ForcedActivations, oracle_inputs, rel_err
from oracle import target_vae_oracle as orc
from tvae_b200.config import HotPathConfig

pytestmark = pytest.mark.gpu

DEV = "cuda"


def small_cfgs():
    return [
        HotPathConfig("t_mnist", C=1, n=24, k=9, p=3, G=8, z=2, O=32, hidden=128),
        HotPathConfig("t_galaxy", C=3, n=16, k=16, p=8, G=4, z=3, O=64, hidden=64, gen_layers=4, n_out=3,
                      likelihood="bernoulli_rgb"),
        HotPathConfig("t_part", C=1, n=20, k=11, p=2, G=16, z=8, O=128, hidden=256, likelihood="gaussian", ctf=True),
        HotPathConfig("t_dsp", C=1, n=16, k=16, p=8, G=8, z=2, O=32, hidden=64, fourier=False, normal_prior_over_r=True),
    ]


def _ops():
    from tvae_b200 import ops
    return ops


@pytest.mark.parametrize("cfg", small_cfgs(), ids=lambda c: c.name)
def test_filter_bank_fwd_bwd(cfg):
    ops = _ops()
    enc, gen, x, y, ctf, nz = oracle_inputs(cfg, 2, dtype=torch.float64, requires_grad=False)
    s = ops.enc_shape(2, cfg.C, cfg.n, cfg.k, cfg.p, cfg.G, cfg.O, cfg.z)
    w = enc.conv1_w.float().to(DEV)
    bank = ops.filter_bank_fwd(s, w).cpu()
    K = cfg.C * cfg.k ** 2
    ref = orc.rotated_filter_bank(enc.conv1_w, cfg.G)            # (O,G,C,1,k,k)
    ref = ref.permute(1, 0, 2, 3, 4, 5).reshape(cfg.G * cfg.O, K)  # row r*O + o
    assert rel_err(bank[:, :K], ref) < 6e-4                        # fp16 rounding of the stored bank (11-bit significand, like tf32)
    assert bank.shape[1] % 64 == 0 and float(bank[:, K:].abs().sum()) == 0.0
    # adjoint: <bank(w), D> == <w, bank^T(D)>
    g = torch.Generator().manual_seed(1)
    D = torch.randn(cfg.G * cfg.O, s.kpad, generator=g)
    dw, db = ops.filter_bank_bwd(s, D.to(DEV))
    wv = enc.conv1_w.clone().requires_grad_(True)
    refb = orc.rotated_filter_bank(wv, cfg.G).permute(1, 0, 2, 3, 4, 5).reshape(cfg.G * cfg.O, K)
    (refb * D[:, :K].double()).sum().backward()
    assert rel_err(dw.cpu(), wv.grad) < 1e-5
    ref_db = D[:, K].view(cfg.G, cfg.O).sum(0)
    assert rel_err(db.cpu(), ref_db) < 1e-5


def _encoder_inputs(cfg, B):
    ops = _ops()
    enc, gen, x, y, ctf, nz = oracle_inputs(cfg, B, dtype=torch.float64, requires_grad=True)
    s = ops.enc_shape(B, cfg.C, cfg.n, cfg.k, cfg.p, cfg.G, cfg.O, cfg.z)
    p_r = ops.rotation_log_prior(cfg.G, cfg.rot_refinement, cfg.normal_prior_over_r, cfg.theta_prior)
    offs = ops.rotation_offsets(cfg.G, cfg.rot_refinement)
    t = lambda v: v.detach().float().to(DEV)
    wh, bh, add = ops.head_tables(t(enc.conv_a_w), t(enc.conv_a_b), t(enc.conv_r_w), t(enc.conv_r_b), t(enc.conv_z_w),
                                  t(enc.conv_z_b), cfg.G, p_r, offs, DEV)
    return ops, enc, y, nz, s, wh, bh, add, t


def _oracle_heads(cfg, enc, y):
    attn, theta, z = orc.encoder_head_maps(y, enc, cfg.G, cfg.p)
    dt = y.dtype
    attn = attn + orc.rotation_log_prior(cfg.G, cfg.rot_refinement, cfg.normal_prior_over_r, cfg.theta_prior, dt)
    offs = orc.rotation_offsets(cfg.G, cfg.rot_refinement, dt)
    theta = torch.stack((theta[:, 0] + offs.view(1, cfg.G, 1, 1), theta[:, 1]), 1)
    B = y.shape[0]
    return torch.cat([attn.unsqueeze(1), theta, z], 1).reshape(B, 3 + 2 * cfg.z, cfg.G, -1)


@pytest.mark.parametrize("cfg", small_cfgs(), ids=lambda c: c.name)
def test_encoder_fwd_bwd(cfg):
    B = 3
    ops, enc, y, nz, s, wh, bh, add, t = _encoder_inputs(cfg, B)
    bank = ops.filter_bank_fwd(s, t(enc.conv1_w))
    x1, h, heads, _ = ops.encoder_fwd(s, t(y), bank, t(enc.conv1_b), t(enc.conv2_w).view(cfg.O, cfg.O), t(enc.conv2_b), wh, bh, add)
    torch.cuda.synchronize()
    ref_heads = _oracle_heads(cfg, enc, y)
    # conv1 activation, internal layout [(b*G + r)*P + pos][O]
    ref_x1 = F.leaky_relu(orc.groupconv_forward(y, enc.conv1_w, enc.conv1_b, cfg.G, cfg.p), 0.01)  # (B,O,G,H,W)
    ref_x1 = ref_x1.permute(0, 2, 3, 4, 1).reshape(-1, cfg.O)
    e1 = rel_err(x1.cpu(), ref_x1.detach())
    eh = rel_err(heads.cpu(), ref_heads.detach())
    print(f"{cfg.name}: conv1 rel err {e1:.2e}, heads rel err {eh:.2e}")
    assert e1 < 3e-3
    assert eh < 3e-3
    # backward against autograd of the oracle (same activation pattern) for a random cotangent on the head maps
    g = torch.Generator().manual_seed(2)
    D = torch.randn(ref_heads.shape, generator=g, dtype=torch.float64)
    d = cfg.Hout
    to_ref = lambda a: a.cpu().double().view(B, cfg.G, d, d, cfg.O).permute(0, 4, 1, 2, 3)
    with ForcedActivations([to_ref(x1), to_ref(h)]):
        forced_heads = _oracle_heads(cfg, enc, y)
    (forced_heads * D).sum().backward()
    dbank, dw2, db2, dwh, dbh = ops.encoder_bwd(s, t(y), t(enc.conv2_w).view(cfg.O, cfg.O), wh, x1, h, D.float().to(DEV))
    dw1, db1 = ops.filter_bank_bwd(s, dbank)
    torch.cuda.synchronize()
    NH = 3 + 2 * cfg.z
    ref_dwh = torch.cat([enc.conv_a_w.grad.view(1, -1), enc.conv_r_w.grad.view(2, -1), enc.conv_z_w.grad.view(NH - 3, -1)], 0)
    ref_dbh = torch.cat([enc.conv_a_b.grad.view(1), enc.conv_r_b.grad.view(2), enc.conv_z_b.grad.view(-1)], 0)
    errs = {
        "dwh": rel_err(dwh.cpu(), ref_dwh), "dbh": rel_err(dbh.cpu(), ref_dbh),
        "dw2": rel_err(dw2.cpu(), enc.conv2_w.grad.view(cfg.O, cfg.O)), "db2": rel_err(db2.cpu(), enc.conv2_b.grad),
        "dw1": rel_err(dw1.cpu(), enc.conv1_w.grad), "db1": rel_err(db1.cpu(), enc.conv1_b.grad),
    }
    print(f"{cfg.name}: encoder grad rel errs {errs}")
    for k, v in errs.items():
        assert v < 5e-3, (k, v)


@pytest.mark.parametrize("cfg", small_cfgs()[:3], ids=lambda c: c.name)
def test_encoder_bwd_wide_dynamic_range(cfg):
    """Head-map gradients span far more than fp16's exponent range in a real step: the attention-logit channel is O(1 / B),
    the theta / z channels are q(t, r) / B times an O(1) factor - 2.6e-8 at cfg5 (150 k cells, B = 256).  They are an fp16 MMA
    operand of the heads backward, so they are scaled by a power of two chosen on the device before the rounding
    (scales[6], enc_bwd_scales_kernel); unscaled, the theta / z rows of dWh flushed to zero.  Cotangent here: channel 0 at
    O(1), every other channel at 1e-8: each row block of dWh must keep its relative accuracy."""
    B = 3
    ops, enc, y, nz, s, wh, bh, add, t = _encoder_inputs(cfg, B)
    bank = ops.filter_bank_fwd(s, t(enc.conv1_w))
    x1, h, heads, _ = ops.encoder_fwd(s, t(y), bank, t(enc.conv1_b), t(enc.conv2_w).view(cfg.O, cfg.O), t(enc.conv2_b), wh, bh, add)
    torch.cuda.synchronize()
    ref_heads = _oracle_heads(cfg, enc, y)
    g = torch.Generator().manual_seed(4)
    D = torch.randn(ref_heads.shape, generator=g, dtype=torch.float64)
    D[:, 1:] *= 1e-8
    d = cfg.Hout
    to_ref = lambda a: a.cpu().double().view(B, cfg.G, d, d, cfg.O).permute(0, 4, 1, 2, 3)
    with ForcedActivations([to_ref(x1), to_ref(h)]):
        forced_heads = _oracle_heads(cfg, enc, y)
    (forced_heads * D).sum().backward()
    dbank, dw2, db2, dwh, dbh = ops.encoder_bwd(s, t(y), t(enc.conv2_w).view(cfg.O, cfg.O), wh, x1, h, D.float().to(DEV))
    torch.cuda.synchronize()
    errs = {"conv_a.weight": rel_err(dwh[0].cpu(), enc.conv_a_w.grad.view(-1)),
            "conv_r.weight": rel_err(dwh[1:3].cpu(), enc.conv_r_w.grad.view(2, -1)),
            "conv_z.weight": rel_err(dwh[3:].cpu(), enc.conv_z_w.grad.view(-1, cfg.O)),
            "conv_z.bias": rel_err(dbh[3:].cpu(), enc.conv_z_b.grad.view(-1))}
    print(f"{cfg.name}: wide-range head gradients, rel errs {errs}")
    for k, v in errs.items():
        assert v < 5e-3, (k, v)


def _enc_out_from_heads(heads, cfg, gumbel, dt):
    B = heads.shape[0]
    d = cfg.Hout
    attn = heads[:, 0].reshape(B, cfg.G, d, d)
    theta = heads[:, 1:3].reshape(B, 2, cfg.G, d, d)
    z = heads[:, 3:].reshape(B, 2 * cfg.z, cfg.G, d, d)
    q = F.log_softmax(attn.reshape(B, -1), 1).view_as(attn)
    a = F.softmax(attn.reshape(B, -1) + gumbel, 1).view_as(attn)
    p_r = orc.rotation_log_prior(cfg.G, cfg.rot_refinement, cfg.normal_prior_over_r, cfg.theta_prior, dt)
    offs = orc.rotation_offsets(cfg.G, cfg.rot_refinement, dt)
    return attn, q, p_r, a, offs, theta, z


@pytest.mark.parametrize("cfg", small_cfgs(), ids=lambda c: c.name)
def test_attention_fwd_bwd(cfg):
    ops = _ops()
    B, d = 4, cfg.Hout
    NH = 3 + 2 * cfg.z
    g = torch.Generator().manual_seed(3)
    heads = (torch.randn(B, NH, cfg.G, d * d, generator=g, dtype=torch.float64) * 0.7).requires_grad_(True)
    from tvae_b200 import synth
    nz = {k: torch.from_numpy(v).double() for k, v in synth.noise(cfg, B, 5).items()}
    x = orc.image_coords(cfg.n, torch.float64)
    enc_out = _enc_out_from_heads(heads, cfg, nz["gumbel"], torch.float64)
    z_b, theta_b, dx, x_t, kl = orc.attention_posterior(enc_out, x, nz["r_z"], nz["r_theta"], cfg.G, cfg.G, cfg.theta_prior)
    spacing = float(x[1, 0] - x[0, 0])
    offs = ops.rotation_offsets(cfg.G, cfg.rot_refinement)
    p_r = ops.rotation_log_prior(cfg.G, cfg.rot_refinement, cfg.normal_prior_over_r, cfg.theta_prior)
    s = ops.attn_shape(B, cfg.G, d, cfg.z, spacing, offs)
    lp = ops.attn_log_prior(s, p_r, DEV)
    hg = heads.detach().float().to(DEV)
    gum, rz, rth = nz["gumbel"].float().to(DEV), nz["r_z"][:, :, 0].float().contiguous().to(DEV), nz["r_theta"].view(B).float().to(DEV)
    out = ops.attn_fwd(s, hg, gum, rz, rth, lp)
    torch.cuda.synchronize()
    assert rel_err(out["zb"].cpu(), z_b.detach()) < 1e-4
    assert rel_err(out["theta_b"].cpu(), theta_b.detach()) < 1e-4
    assert rel_err(out["dx"].cpu(), dx.detach().view(B, 2)) < 1e-4
    assert abs(float(out["kl"].double().mean().cpu()) - float(kl.detach())) < 1e-4 * abs(float(kl.detach()))
    # module-interface tail
    q, a = ops.attn_softmax_pair(hg, gum)
    assert rel_err(q.cpu(), enc_out[1].detach().reshape(B, -1)) < 1e-5
    assert rel_err(a.cpu(), enc_out[3].detach().reshape(B, -1)) < 1e-4
    # backward: random cotangents on (z_b, theta_b, dx) and weight g_kl on the per-image KL
    gz = torch.randn(B, cfg.z, generator=g, dtype=torch.float64)
    gt = torch.randn(B, generator=g, dtype=torch.float64)
    gd = torch.randn(B, 2, generator=g, dtype=torch.float64)
    loss = (z_b * gz).sum() + (theta_b * gt).sum() + (dx.view(B, 2) * gd).sum() + kl
    loss.backward()
    gkl = torch.full((1,), 1.0 / B, device=DEV)
    dh = ops.attn_bwd(s, hg, gum, rz, rth, lp, out, gz.float().to(DEV), gt.float().to(DEV), gd.float().to(DEV), gkl)
    torch.cuda.synchronize()
    e = rel_err(dh.cpu(), heads.grad)
    print(f"{cfg.name}: attention d_heads rel err {e:.2e}")
    assert e < 2e-4
    # get_latent
    zc, th, dxl, am = ops.get_latent(s, hg)
    attn = heads.detach()[:, 0].reshape(B, -1)
    ind = attn.argmax(1)
    assert torch.equal(am.cpu().long(), ind)
    ar = torch.arange(B)
    hv = heads.detach().reshape(B, NH, -1)
    ref_zc = torch.cat([hv[ar, 3:3 + cfg.z, ind], torch.exp(hv[ar, 3 + cfg.z:, ind])], 1)
    assert rel_err(zc.cpu(), ref_zc) < 1e-5
    assert rel_err(th.cpu().view(B), hv[ar, 1, ind]) < 1e-6
    ref_dx = F.softmax(attn, 1).view(B, cfg.G, d * d).sum(1) @ orc.translation_grid(d, x[1, 0] - x[0, 0], torch.float64)
    assert rel_err(dxl.cpu(), ref_dx) < 1e-4


@pytest.mark.parametrize("cfg", small_cfgs(), ids=lambda c: c.name)
@pytest.mark.parametrize("explicit_coords", [False, True])
def test_generator_fwd_bwd(cfg, explicit_coords):
    ops = _ops()
    B = 3
    enc, gen, x, y, ctf, nz = oracle_inputs(cfg, B, dtype=torch.float64, requires_grad=True)
    g = torch.Generator().manual_seed(4)
    theta = (torch.randn(B, generator=g, dtype=torch.float64)).requires_grad_(True)
    dx = (torch.randn(B, 2, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True)
    zb = torch.randn(B, cfg.z, generator=g, dtype=torch.float64).requires_grad_(True)
    c, sn = torch.cos(theta), torch.sin(theta)
    rot = torch.stack([torch.stack([c, sn], 1), torch.stack([-sn, c], 1)], 1)
    xt = torch.bmm(x.expand(B, -1, 2) - dx.unsqueeze(1), rot)
    if explicit_coords:
        xt = xt.detach().requires_grad_(True)
    y_ref = orc.generator_forward(xt, zb, gen)
    t = lambda v: v.detach().float().to(DEV)
    wf = None if gen.fourier_w is None else t(gen.fourier_w / torch.tensor(gen.sigma, dtype=torch.float32).double())
    if wf is not None:
        # the product divides in fp32 like the reference (models.py:57)
        wf = (gen.fourier_w.float() / torch.tensor(gen.sigma, dtype=torch.float32)).to(DEV)
    gw = ops.GenWeights(wf, None if gen.fourier_b is None else t(gen.fourier_b), t(gen.coord_w), t(gen.coord_b), t(gen.latent_w),
                        [t(w) for w in gen.hidden_w], [t(b) for b in gen.hidden_b], t(gen.out_w), t(gen.out_b))
    N = cfg.n ** 2
    s = ops.gen_shape(B, N, gw, cfg.z)
    if explicit_coords:
        xin, th_in, dx_in = t(xt).reshape(B * N, 2), None, None
    else:
        xin, th_in, dx_in = t(x), t(theta), t(dx)
    y_hat, saved = ops.generator_fwd(s, gw, xin, th_in, dx_in, t(zb))
    torch.cuda.synchronize()
    e = rel_err(y_hat.cpu().view(B, N, -1), y_ref.detach())
    print(f"{cfg.name} explicit={explicit_coords}: generator y_hat rel err {e:.2e}")
    assert e < 5e-3
    D = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    with ForcedActivations([a.cpu().double() for a in saved["acts"]]):
        y_forced = orc.generator_forward(xt, zb, gen)
    (y_forced * D).sum().backward()
    out = ops.generator_bwd(s, gw, xin, th_in, dx_in, t(zb), saved, y_hat, D.float().to(DEV).reshape(B * N, -1))
    torch.cuda.synchronize()
    errs = {
        "dw1": rel_err(out["dw1"].cpu(), gen.coord_w.grad), "db1": rel_err(out["db1"].cpu(), gen.coord_b.grad),
        "dwz": rel_err(out["dwz"].cpu(), gen.latent_w.grad), "dwout": rel_err(out["dwout"].cpu(), gen.out_w.grad),
        "dbout": rel_err(out["dbout"].cpu(), gen.out_b.grad), "d_z": rel_err(out["d_z"].cpu(), zb.grad),
    }
    for i, (w, b) in enumerate(zip(gen.hidden_w, gen.hidden_b)):
        errs[f"dwh{i}"] = rel_err(out["dwh"][i].cpu(), w.grad)
        errs[f"dbh{i}"] = rel_err(out["dbh"][i].cpu(), b.grad)
    if explicit_coords:
        errs["dxp"] = rel_err(out["dxp"].cpu().view(B, N, 2), xt.grad)
    else:
        errs["d_theta"] = rel_err(out["d_theta"].cpu(), theta.grad)
        errs["d_dx"] = rel_err(out["d_dx"].cpu(), dx.grad)
    print(f"{cfg.name} explicit={explicit_coords}: generator grad rel errs " + ", ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    for k, v in errs.items():
        assert v < 5e-3, (k, v)


@pytest.mark.parametrize("layers,explicit_coords", [(2, False), (3, False), (2, True)])
def test_generator_coord_fused_matches_unfused(layers, explicit_coords):
    """Generators without Fourier features (cfg2): the coordinate layer fused into the first hidden layer's GEMM (a0 never
    stored; hidden weight gradient regenerates it, input gradient masks with one bit per element) against the unfused
    kernels of the same library (development knob 1 switches the fusion off) and against the fp64 oracle's forward.
    n = 13 puts image boundaries inside 128-row tiles (169 pixels per image); layers = 3 leaves the projection to the
    second hidden layer."""
    ops = _ops()
    cfg = HotPathConfig("t_fused", C=1, n=13, k=7, p=3, G=4, z=3, O=32, hidden=256, gen_layers=layers, fourier=False)
    B = 5
    enc, gen, x, y, ctf, nz = oracle_inputs(cfg, B, dtype=torch.float64)
    g = torch.Generator().manual_seed(11)
    theta = torch.randn(B, generator=g, dtype=torch.float64)
    dx = torch.randn(B, 2, generator=g, dtype=torch.float64) * 0.1
    zb = torch.randn(B, cfg.z, generator=g, dtype=torch.float64)
    c, sn = torch.cos(theta), torch.sin(theta)
    rot = torch.stack([torch.stack([c, sn], 1), torch.stack([-sn, c], 1)], 1)
    xt = torch.bmm(x.expand(B, -1, 2) - dx.unsqueeze(1), rot)
    y_ref = orc.generator_forward(xt, zb, gen)
    t = lambda v: v.detach().float().to(DEV)
    gw = ops.GenWeights(None, None, t(gen.coord_w), t(gen.coord_b), t(gen.latent_w), [t(w) for w in gen.hidden_w],
                        [t(b) for b in gen.hidden_b], t(gen.out_w), t(gen.out_b))
    N = cfg.n ** 2
    s = ops.gen_shape(B, N, gw, cfg.z)
    xin, th_in, dx_in = (t(xt).reshape(B * N, 2), None, None) if explicit_coords else (t(x), t(theta), t(dx))
    D = torch.randn(B * N, cfg.n_out, generator=g).to(DEV)
    res = {}
    try:
        for fused in (True, False):
            ops.L().tvae_test_set_knob(1, 0 if fused else 1)
            y_hat, saved = ops.generator_fwd(s, gw, xin, th_in, dx_in, t(zb))
            out = ops.generator_bwd(s, gw, xin, th_in, dx_in, t(zb), saved, y_hat, D)
            torch.cuda.synchronize()
            res[fused] = (y_hat.clone(), {k: v.clone() for k, v in out.items() if k != "flat" and v is not None})
    finally:
        ops.L().tvae_test_set_knob(1, 0)
    assert rel_err(res[True][0].cpu().view(B, N, -1), y_ref) < 5e-3
    assert rel_err(res[True][0].cpu(), res[False][0].cpu()) < 2e-3
    for k, v in res[True][1].items():
        if explicit_coords and k in ("d_theta", "d_dx"):          # not computed with explicit coordinates (dxp is the gradient)
            continue
        e = rel_err(v.cpu(), res[False][1][k].cpu())
        print(f"fused vs unfused {k}: {e:.1e}")
        # a unit whose pre-activation rounds to the other side of zero in the two arithmetic orders flips its derivative
        assert e < 2e-2, (k, e)


def test_bernoulli_and_gaussian():
    ops = _ops()
    g = torch.Generator().manual_seed(6)
    B, n = 3, 20
    yh = torch.randn(B, n * n, generator=g, dtype=torch.float64).requires_grad_(True)
    y = torch.rand(B, n * n, generator=g, dtype=torch.float64)
    ll_ref = -(F.binary_cross_entropy_with_logits(yh, y, reduction="none")).sum(1)
    gsc = torch.tensor([-1.0 / B], device=DEV)
    ll, d = ops.bernoulli(yh.detach().float().to(DEV), y.float().to(DEV), gsc)
    (ll_ref.sum() * (-1.0 / B)).backward()
    assert rel_err(ll.cpu(), ll_ref.detach()) < 1e-5
    assert rel_err(d.cpu(), yh.grad) < 1e-5
    # gaussian + CTF (+ mask)
    from tvae_b200 import synth
    import numpy as np
    # CUDA-core correlation kernel (fp32, 1e-4) and the tensor-core banded-Toeplitz GEMM (fp16 operands: 2^-11 relative
    # per element -> 2e-3), at a toy size and at the particle-stack size n = 128
    for n, radii in ((20, (0, 6)), (128, (0,))):
        ctf = torch.from_numpy(synth.ctf_kernels(np.random.default_rng(0), B, n)).double()
        for radius in radii:
            yh = torch.randn(B, n * n, generator=g, dtype=torch.float64).requires_grad_(True)
            yy = torch.randn(B, n * n, generator=g, dtype=torch.float64)
            dx = torch.randn(B, 1, 2, generator=g, dtype=torch.float64) * 0.1
            s = torch.tensor(2.0 / (n - 1), dtype=torch.float32)
            mask = orc.particle_mask(dx, s, n, radius) if radius else None
            mu = F.conv2d(yh.view(1, B, n, n), ctf, padding=(n - 1) // 2, groups=B).view(B, -1)
            diff = mu - yy
            if mask is not None:
                diff = torch.where(mask, diff, torch.zeros_like(diff))
            ll_ref = -0.5 * (diff ** 2).sum(1)
            (ll_ref.sum() * (-1.0 / B)).backward()
            for gemm, tol in ((False, 1e-4), (True, 2e-3)):
                ll, d, mu_k = ops.gaussian(yh.detach().float().to(DEV), yy.float().to(DEV), n, ctf.float().to(DEV),
                                           dx.view(B, 2).float().to(DEV), float(s), radius, gsc, use_ctf_gemm=gemm)
                torch.cuda.synchronize()
                e = (rel_err(mu_k.cpu(), mu.detach()), rel_err(ll.cpu(), ll_ref.detach()), rel_err(d.cpu(), yh.grad))
                print(f"gaussian+CTF n={n} radius={radius} gemm={gemm}: rel err mu {e[0]:.2e} ll {e[1]:.2e} d_yhat {e[2]:.2e}")
                assert max(e) < tol, (n, radius, gemm, e)
                # backward-only call that reuses the forward's mu
                _, d2, _ = ops.gaussian(yh.detach().float().to(DEV), yy.float().to(DEV), n, ctf.float().to(DEV),
                                        dx.view(B, 2).float().to(DEV), float(s), radius, gsc, mu=mu_k, use_ctf_gemm=gemm)
                assert rel_err(d2.cpu(), d.cpu()) < 1e-6


@pytest.mark.parametrize("n,m", [(16, 9), (16, 31), (16, 41), (20, 39), (64, 127)])
def test_gaussian_ctf_filter_of_any_odd_size(n, m):
    """train_particles.py:298-302 applies whatever odd-sized (B,1,m,m) filter it is given with padding m // 2 (with
    --crop the filters keep the uncropped micrograph's size, :543-547): smaller than, larger than, and (m > 2n - 1) far
    larger than the image, against F.conv2d in fp64."""
    ops = _ops()
    g = torch.Generator().manual_seed(n * 1000 + m)
    B = 3
    ctf = torch.randn(B, 1, m, m, generator=g, dtype=torch.float64) / m
    yh = torch.randn(B, n * n, generator=g, dtype=torch.float64).requires_grad_(True)
    yy = torch.randn(B, n * n, generator=g, dtype=torch.float64)
    gsc = torch.tensor([-1.0 / B], device=DEV)
    mu = F.conv2d(yh.view(1, B, n, n), ctf, padding=m // 2, groups=B).view(B, -1)
    ll_ref = -0.5 * ((mu - yy) ** 2).sum(1)
    (ll_ref.sum() * (-1.0 / B)).backward()
    ll, d, mu_k = ops.gaussian(yh.detach().float().to(DEV), yy.float().to(DEV), n, ctf.float().to(DEV), None, 1.0, 0, gsc)
    torch.cuda.synchronize()
    e = (rel_err(mu_k.cpu(), mu.detach()), rel_err(ll.cpu(), ll_ref.detach()), rel_err(d.cpu(), yh.grad))
    assert max(e) < 1e-4, (n, m, e)


def test_gaussian_ctf_rejects_bad_filters():
    ops = _ops()
    B, n = 2, 16
    yh = torch.zeros(B, n * n, device=DEV)
    for shape in ((B, 1, 10, 10), (B + 1, 1, 15, 15), (B, 1, 15, 13)):      # even size, wrong batch, not square
        with pytest.raises(ValueError):
            ops.gaussian(yh, yh, n, torch.zeros(*shape, device=DEV))
