"""torch.autograd.Functions over the C ABI.

  GroupConvFn        GroupConv.forward                         (models.py:202-225)
  EncoderHeadsFn     encoder up to the attention / theta / z maps (models.py:354-358, 382, 390-399)
  GeneratorFn        SpatialGenerator.forward                   (models.py:95-123)
  FusedStepFn        the whole attention/attention(+offsets) branch of eval_minibatch
                     (train_mnist.py:187-294, train_particles.py:186-343) in one node

Backward runs on the autograd engine thread on the current stream, like the reference's backward.
"""
from __future__ import annotations

import functools
import math
from dataclasses import dataclass
from typing import Callable, Optional

import torch

from . import ops


class nvtx_range:
    """NVTX range around one subsystem of the step (visible in nsys / ncu --nvtx; a no-op without a profiler attached)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        torch.cuda.nvtx.range_pop()


def on_tensor_device(fn):
    """The library launches on the CURRENT CUDA device (stream, SM count, shared-memory opt-in): run `fn` with the
    device of its first CUDA tensor argument current, so a model on cuda:1 works while cuda:0 is current.  (Backward
    passes need nothing: the autograd engine already switches to the device of the incoming gradients.)"""
    @functools.wraps(fn)
    def wrapped(*args, **kw):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kw)
        return fn(*args, **kw)
    return wrapped


# ------------------------------------------------------------------------------------------------ GroupConv
class GroupConvFn(torch.autograd.Function):
    @staticmethod
    @on_tensor_device
    def forward(ctx, y, weight, bias, G, padding):
        B, n = y.shape[0], y.shape[-1]
        O, C, _, k, _ = weight.shape
        s = ops.enc_shape(B, C, n, k, padding, G, O, 1)
        yc = ops.f32(y).reshape(B, C, n, n)
        bank = ops.filter_bank_fwd(s, weight)
        d = n + 2 * padding - k + 1
        out = ops.empty(B * G * d * d, O, device=y.device)
        ops.check(ops.L().tvae_groupconv_fwd(ops.byref(s), ops.ptr(yc), ops.ptr(bank),
                                             ops.ptr(None if bias is None else ops.f32(bias)), ops.ptr(out),
                                             ops.stream_ptr()), "tvae_groupconv_fwd")
        ctx.s, ctx.has_bias = s, bias is not None
        ctx.weight, ctx.yshape = weight.detach(), y.shape
        ctx.save_for_backward(yc)
        # internal [(b,r,pos)][o] -> reference (B,O,G,H',W')
        return out.view(B, G, d, d, O).permute(0, 4, 1, 2, 3)

    @staticmethod
    def backward(ctx, g):
        (yc,) = ctx.saved_tensors
        s = ctx.s
        gi = g.permute(0, 2, 3, 4, 1).contiguous().view(-1, s.O).float()
        dy = None
        if ctx.needs_input_grad[0]:
            # off the hot path (the training step's image is data): CUDA-core transposed convolution with the fp32 bank
            dy = ops.empty(s.B, s.C, s.n, s.n, device=g.device)
            bank32 = ops.empty(s.G * s.O, s.C * s.k * s.k, device=g.device)
            ops.check(ops.L().tvae_groupconv_dgrad(ops.byref(s), ops.ptr(ops.f32(ctx.weight)), ops.ptr(gi), ops.ptr(bank32), ops.ptr(dy),
                                                   ops.stream_ptr()), "tvae_groupconv_dgrad")
            dy = dy.view(ctx.yshape)
        dbank = ops.empty(s.G * s.O, s.kpad, device=g.device)
        gi16 = ops.half(*gi.shape, device=g.device)
        scales = ops.empty(8, device=g.device)
        ops.check(ops.L().tvae_groupconv_wgrad(ops.byref(s), ops.ptr(yc), ops.ptr(gi), ops.ptr(gi16), ops.ptr(scales), ops.ptr(dbank),
                                               ops.stream_ptr()),
                  "tvae_groupconv_wgrad")
        dw, db = ops.filter_bank_bwd(s, dbank)
        return dy, dw, (db if ctx.has_bias else None), None, None


# ------------------------------------------------------------------------------------------------ encoder heads
@dataclass
class EncoderSpec:
    G: int
    padding: int
    z: int
    rot_refinement: bool
    normal_prior_over_r: bool
    theta_prior: float
    # std of the Normal prior on theta given (t, r): pi / G for the attention/attention branch (train_mnist.py:269-272);
    # the attention/unimodal branch uses eval_minibatch's theta_prior argument instead (train_mnist.py:171)
    theta_prior_std: Optional[float] = None
    act: int = ops.ACT_LEAKYRELU       # --activation of the encoder (ops.ACT_*)
    # rotation pooling through fc_r between conv1 and conv2 (attention/unimodal encoder with groupconv > 0,
    # models.py:301-304): the group conv keeps G rotations, everything after it has one rotation slot
    pool: bool = False

    @property
    def attn_G(self) -> int:
        return 1 if self.pool else self.G

    @property
    def n_params(self) -> int:
        return 12 if self.pool else 10

    def tables(self):
        p_r = ops.rotation_log_prior(self.attn_G, self.rot_refinement, self.normal_prior_over_r, self.theta_prior)
        offs = ops.rotation_offsets(self.attn_G, self.rot_refinement)
        return p_r, offs


ENC_PARAM_NAMES = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv_a.weight", "conv_a.bias",
                   "conv_r.weight", "conv_r.bias", "conv_z.weight", "conv_z.bias"]


def _encoder_forward(spec: EncoderSpec, y, w1, b1, w2, b2, wa, ba, wr, br, wz, bz, fcw=None, fcb=None, keep_h=True):
    B, n = y.shape[0], y.shape[-1]
    if w1.dim() == 4:      # nn.Conv2d weight (O,C,k,k) of the groupconv = 0 encoder: one rotation, no rotation axis
        w1 = w1.unsqueeze(2)
    O, C, _, k, _ = w1.shape
    s = ops.enc_shape(B, C, n, k, spec.padding, spec.G, O, spec.z, spec.act)
    p_r, offs = spec.tables()
    yc = ops.f32(y).reshape(B, C, n, n)
    wh, bh, add = ops.head_tables(wa, ba, wr, br, wz, bz, spec.attn_G, p_r, offs, y.device)
    bank = ops.filter_bank_fwd(s, w1)
    w2m = ops.f32(w2).reshape(O, O)
    x1, h, heads, xp = ops.encoder_fwd(s, yc, bank, b1, w2m, b2, wh, bh, add, keep_h, (fcw, fcb) if spec.pool else None)
    return s, yc, w2m, wh, x1, h, heads, xp


@on_tensor_device
def _encoder_forward_inference(spec, y, *params):
    return _encoder_forward(spec, y, *params, keep_h=False)


def encoder_heads_inference(spec: EncoderSpec, y, *params):
    """heads (B, 3+2z, G, H', W') without autograd and without keeping the hidden map (clustering_*.get_latent)."""
    with torch.no_grad():
        s, _, _, _, _, _, heads, _ = _encoder_forward_inference(spec, y, *params)
    d = s.n + 2 * s.p - s.k + 1
    return heads.view(s.B, 3 + 2 * spec.z, spec.attn_G, d, d)


@on_tensor_device
def refine_argmax(spec: EncoderSpec, y, heads, *params, rel_tol=ops.REFINE_REL_TOL):
    """fp32-accurate argmax (r, t) and the z / theta values there: the fast head maps select the candidate cells, the
    logit chain is re-evaluated exactly at them (ops.refine_argmax).  heads: (B, NH, G2, H', W') from
    encoder_heads_inference on the same images and parameters."""
    w1, b1, w2, b2, wa, ba, wr, br, wz, bz = params[:10]
    if w1.dim() == 4:
        w1 = w1.unsqueeze(2)
    O, C, _, k, _ = w1.shape
    B, n = y.shape[0], y.shape[-1]
    s = ops.enc_shape(B, C, n, k, spec.padding, spec.G, O, spec.z, spec.act)
    p_r, offs = spec.tables()
    wh, bh, add = ops.head_tables(wa, ba, wr, br, wz, bz, spec.attn_G, p_r, offs, y.device)
    with torch.no_grad():
        return ops.refine_argmax(s, ops.f32(y).reshape(B, C, n, n), w1, b1, ops.f32(w2).reshape(O, O), b2, wh, bh, add,
                                 heads.reshape(B, heads.shape[1], spec.attn_G, -1).contiguous(),
                                 pool=(params[10], params[11]) if spec.pool else None, rel_tol=rel_tol)


def _encoder_backward(s, spec: EncoderSpec, yc, w2m, wh, x1, h, d_heads, shapes=None, fcw=None, xp=None):
    """-> (grads, flat): grads in ENC_PARAM_NAMES order [+ fc_r.weight, fc_r.bias with rotation pooling] (Conv3d shapes, or
    reshaped to `shapes` = the parameters' own shapes), all of them views of ONE flat bucket `flat` the kernels wrote
    into (the data-parallel all-reduce runs on it in place)."""
    O, z, NH, dev = s.O, spec.z, 3 + 2 * spec.z, yc.device
    bucket = [(O, s.C, 1, s.k, s.k), (O,), (O, O), (O,), (NH, O), (NH,)] + ([(s.G,), (1,)] if spec.pool else [])
    flat, v = ops.flat_views(bucket, dev)
    res = ops.encoder_bwd(s, yc, w2m, wh, x1, h, d_heads, (fcw, xp) if spec.pool else None, out=v[2:])
    dbank = res[0]
    ops.filter_bank_bwd(s, dbank, out=(v[0], v[1]))
    dw1, db1, dw2, db2, dwh, dbh = v[:6]
    sh = (O, 1, 1, 1)
    grads = [dw1, db1, dw2.view(O, O, 1, 1, 1), db2,
             dwh[0:1].reshape(1, *sh), dbh[0:1], dwh[1:3].reshape(2, *sh), dbh[1:3],
             dwh[3:].reshape(2 * z, *sh), dbh[3:]] + ([v[6].reshape(1, -1), v[7]] if spec.pool else [])
    if shapes is not None:
        grads = [g.reshape(sh_) for g, sh_ in zip(grads, shapes)]
    return grads, flat


class EncoderHeadsFn(torch.autograd.Function):
    """heads (B, 3+2z, G, H', W'): attn(+p_r), theta(+offsets), z stacked on dim 1."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, spec, y, *params):
        s, yc, w2m, wh, x1, h, heads, xp = _encoder_forward(spec, y, *params)
        ctx.spec, ctx.s = spec, s
        ctx.shapes = [p.shape for p in params]
        ctx.save_for_backward(yc, w2m, wh, x1, h, *((params[10], xp) if spec.pool else ()))
        d = s.n + 2 * s.p - s.k + 1
        return heads.view(s.B, 3 + 2 * spec.z, spec.attn_G, d, d)

    @staticmethod
    def backward(ctx, g):
        yc, w2m, wh, x1, h, *pool = ctx.saved_tensors
        grads, _ = _encoder_backward(ctx.s, ctx.spec, yc, w2m, wh, x1, h, g.contiguous(), ctx.shapes, *pool)
        return (None, None, *grads)


# ------------------------------------------------------------------------------------------------ generator
GEN_FIXED = ["coord_linear.weight", "coord_linear.bias", "latent_linear.weight"]


@functools.lru_cache(maxsize=16)
def _identity(n: int, device: str) -> torch.Tensor:
    return torch.eye(n, dtype=torch.float32, device=device)


@functools.lru_cache(maxsize=64)
def _sigma_tensor(sigma: float, device: str) -> torch.Tensor:
    """sigma as an fp32 device scalar, created ONCE per (value, device): torch.tensor(..., device=cuda) is a pageable
    host-to-device copy that waits for the stream - every step paid 0.6 ms of CPU-GPU serialisation for it (cfg1 was
    launch-bound by it: 2.95 ms of host time per 3.0 ms step)."""
    return torch.tensor(sigma, dtype=torch.float32, device=device)


def _gen_weights(fourier_w, fourier_b, sigma, w1, b1, wz, hidden, wout, bout, resid=False):
    """resid: the hidden layers are ResidLinear modules, act(W x + b + x) = act((W + I) x + b) (models.py:29-30; the
    activation comes after the residual add).  They run as plain layers with the effective weight W + I: the forward,
    the input gradient dpre (W + I) and the fp16 gradient-scale bounds all see W + I, and the weight gradient
    dpre^T x is the gradient w.r.t. W as it stands."""
    wf = None
    if fourier_w is not None:
        # F.linear(x, weight / sigma, bias) with sigma an fp32 tensor (models.py:40,57)
        wf = (ops.f32(fourier_w) / _sigma_tensor(float(sigma), str(fourier_w.device))).contiguous()
    hw = [hidden[i] for i in range(0, len(hidden), 2)]
    hb = [hidden[i] for i in range(1, len(hidden), 2)]
    if resid:
        hw = [ops.f32(w) + _identity(w.shape[0], str(w.device)) for w in hw]
    return ops.GenWeights(wf, None if fourier_b is None else ops.f32(fourier_b), w1, b1, wz, hw, hb, wout, bout)


def _gen_param_grads(out, L):
    grads = [out["dw1"], out["db1"], out["dwz"]]
    for i in range(L):
        grads += [out["dwh"][i], out["dbh"][i]]
    grads += [out["dwout"], out["dbout"]]
    return grads


class GeneratorFn(torch.autograd.Function):
    """y_hat (B,N,n_out) from explicit coordinates x (B,N,2) and z (B,zdim).
    params = coord_linear.weight, coord_linear.bias, latent_linear.weight, (hidden w, b)*, out w, out b."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, fourier_w, fourier_b, sigma, x, z, *params):
        resid, act = False, ops.ACT_LEAKYRELU
        if isinstance(sigma, tuple):       # (sigma, resid, act) from SpatialGenerator
            sigma, resid, act = sigma
        w1, b1, wz = params[:3]
        hidden, (wout, bout) = params[3:-2], params[-2:]
        gw = _gen_weights(fourier_w, fourier_b, sigma, w1, b1, wz, hidden, wout, bout, resid)
        B, N = x.shape[0], x.shape[1]
        s = ops.gen_shape(B, N, gw, z.shape[1], act)
        xc, zc = ops.f32(x).reshape(B * N, 2), ops.f32(z)
        y_hat, saved = ops.generator_fwd(s, gw, xc, None, None, zc)
        ctx.s, ctx.gw, ctx.saved = s, gw, saved
        ctx.save_for_backward(xc, zc, y_hat)
        return y_hat.view(B, N, -1)

    @staticmethod
    def backward(ctx, g):
        xc, zc, y_hat = ctx.saved_tensors
        s = ctx.s
        out = ops.generator_bwd(s, ctx.gw, xc, None, None, zc, ctx.saved, y_hat, g.reshape(s.B * s.N, -1).contiguous())
        ctx.saved = None
        return (None, None, None, out["dxp"].view(s.B, s.N, 2), out["d_z"], *_gen_param_grads(out, s.L))


# ------------------------------------------------------------------------------------------------ standalone module interfaces
class FourierEmbedFn(torch.autograd.Function):
    """RandomFourierEmbedding2d.forward (models.py:53-58): cos(x (W / sigma)^T + b) for x (..., 2)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, x, weight, bias, sigma):
        w = (ops.f32(weight) / _sigma_tensor(float(sigma), str(weight.device))).contiguous()
        xc = ops.f32(x).reshape(-1, 2)
        out = ops.fourier_embed_fwd(xc, w, ops.f32(bias))
        ctx.save_for_backward(xc, w, ops.f32(bias))
        ctx.xshape = x.shape
        return out.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, g):
        xc, w, b = ctx.saved_tensors
        dx = ops.fourier_embed_bwd(xc, w, b, ops.f32(g).reshape(xc.shape[0], -1)) if ctx.needs_input_grad[0] else None
        return (None if dx is None else dx.view(ctx.xshape)), None, None, None


class LinearActFn(torch.autograd.Function):
    """act(x W^T + b [+ x]) - ResidLinear.forward (models.py:29-30; the activation follows the residual add, so the layer is a
    Linear with the effective weight W + I) on the tensor-core LinearNT / LinearTN kernels."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, x, weight, bias, resid, act):
        xc = ops.f32(x).reshape(-1, x.shape[-1])
        y, x16 = ops.linear_act_fwd(xc, ops.f32(weight), ops.f32(bias), resid, act)
        ctx.save_for_backward(x16, ops.f32(weight), y)
        ctx.resid, ctx.act, ctx.xshape = resid, act, x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, g):
        x16, w, y = ctx.saved_tensors
        dx, dw, db = ops.linear_act_bwd(x16, w, y, ops.f32(g).reshape(y.shape), ctx.resid, ctx.act, ctx.needs_input_grad[0])
        return (None if dx is None else dx.view(ctx.xshape)), dw, db, None, None


class SoftmaxPairFn(torch.autograd.Function):
    """(q_t_r, a_sampled) = (log_softmax(attn), softmax(attn + gumbel)) over all (r, t) cells (models.py:383-388):
    the module-interface tail of the encoders, one kernel forward, one backward."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, attn, gumbel):
        B = attn.shape[0]
        a2 = ops.f32(attn).reshape(B, 1, 1, -1)
        q, a = ops.attn_softmax_pair(a2, ops.f32(gumbel).reshape(B, -1))
        ctx.save_for_backward(q, a)
        ctx.shape = attn.shape
        return q.view(attn.shape), a.view(attn.shape)

    @staticmethod
    def backward(ctx, dq, da):
        q, a = ctx.saved_tensors
        B = q.shape[0]
        d = ops.attn_softmax_pair_bwd(q, a, None if dq is None else ops.f32(dq).reshape(B, -1),
                                      None if da is None else ops.f32(da).reshape(B, -1))
        return d.view(ctx.shape), None


def gumbel_noise(shape, device):
    """the draw of F.gumbel_softmax (models.py:387): -log(Exp(1)) per cell, from torch's generator on the device"""
    return -torch.empty(shape, device=device, dtype=torch.float32).exponential_().log()


# ------------------------------------------------------------------------------------------------ fused step
_spacing_cache: dict = {}      # id(tensor) -> (weakref to the tensor, its _version, spacing)


def pixel_spacing(x_coord: torch.Tensor) -> float:
    """btw_pixels_space = x[1,0] - x[0,0] (train_mnist.py:30).  The reference syncs for it every step; the coordinate
    grid is constant, so it is read once per grid: cached on the IDENTITY and version counter of the caller's own tensor
    (never on a data pointer - the caching allocator hands freed addresses to other grids), so a new or an in-place
    rescaled grid is read again.  Call it on the tensor the trainer holds, not on a per-step `.to(device)` copy."""
    import weakref
    hit = _spacing_cache.get(id(x_coord))
    if hit is not None and hit[0]() is x_coord and hit[1] == x_coord._version:
        return hit[2]
    v = float((x_coord[1, 0] - x_coord[0, 0]).float().cpu())
    if len(_spacing_cache) > 64:
        for k in [k for k, e in _spacing_cache.items() if e[0]() is None]:
            del _spacing_cache[k]
        if len(_spacing_cache) > 64:
            _spacing_cache.clear()
    _spacing_cache[id(x_coord)] = (weakref.ref(x_coord), x_coord._version, v)
    return v


@dataclass
class StepSpec:
    enc: EncoderSpec
    sigma: float
    likelihood: str = "bernoulli"      # bernoulli | gaussian | gaussian_fit_noise
    mask_radius: int = 0
    n_gen_hidden: int = 1
    gen_resid: bool = False            # --generator-resid-layers: hidden layers are ResidLinear (models.py:22-30)
    gen_act: int = ops.ACT_LEAKYRELU   # --activation of the generator (ops.ACT_*)
    # optional data-parallel gradient synchroniser (tvae_b200.dp.GradSync): bucket 0 (generator) is started as
    # soon as the generator backward has been issued, so its all-reduce overlaps the encoder backward.
    sync: Optional[object] = None
    spacing: Optional[float] = None    # pixel_spacing(x_coord) of the caller's grid tensor (None: read from x_coord here)


class FusedStepFn(torch.autograd.Function):
    """(elbo, log_p_x_g_z, kl_div) of eval_minibatch's attention/attention(+offsets) branch.

    inputs: spec, x_coord (N,2), y (B,C,n,n), ctf or None, gumbel (B,L), r_z (B,z), r_theta (B),
            fourier_w, fourier_b, then the encoder params (ENC_PARAM_NAMES; + fc_r.weight, fc_r.bias with rotation
            pooling) and the generator params.
    """

    @staticmethod
    @on_tensor_device
    def forward(ctx, spec: StepSpec, x_coord, y, ctf, gumbel, r_z, r_theta, fourier_w, fourier_b, *params):
        es = spec.enc
        enc_params, gen_params = params[:es.n_params], params[es.n_params:]
        with nvtx_range("tvae.encoder_fwd"):
            s, yc, w2m, wh, x1, h, heads, xp = _encoder_forward(es, y, *enc_params)
        B, n = s.B, s.n
        d = s.n + 2 * s.p - s.k + 1
        xc = ops.f32(x_coord)
        spacing = spec.spacing if spec.spacing is not None else pixel_spacing(x_coord)
        p_r, offs = es.tables()
        ashape = ops.attn_shape(B, es.attn_G, d, es.z, spacing, offs, es.theta_prior_std)
        log_prior = ops.attn_log_prior(ashape, p_r, y.device)
        gum, rz, rth = ops.f32(gumbel), ops.f32(r_z).reshape(B, es.z), ops.f32(r_theta).reshape(B)
        with nvtx_range("tvae.attention_fwd"):
            att = ops.attn_fwd(ashape, heads, gum, rz, rth, log_prior)

        w1, b1, wz = gen_params[:3]
        hidden, (wout, bout) = gen_params[3:-2], gen_params[-2:]
        gw = _gen_weights(fourier_w, fourier_b, spec.sigma, w1, b1, wz, hidden, wout, bout, spec.gen_resid)
        N = xc.shape[0]
        gs = ops.gen_shape(B, N, gw, es.z, spec.gen_act)
        with nvtx_range("tvae.generator_fwd"):
            y_hat, gsaved = ops.generator_fwd(gs, gw, xc, att["theta_b"], att["dx"], att["zb"])

        yflat = yc.reshape(B, -1)
        ctfc = None if ctf is None else ops.f32(ctf)
        if spec.likelihood == "bernoulli":
            ll, _ = ops.bernoulli(y_hat, yflat)
        elif spec.likelihood == "gaussian_fit_noise":
            ll, _ = ops.gaussian_fit_noise(y_hat, yflat)
        else:
            ll, _, mu = ops.gaussian(y_hat, yflat, n, ctfc, att["dx"], spacing, spec.mask_radius)
            ctx.mu = mu
        log_p = ll.mean()
        kl = att["kl"].mean()
        elbo = log_p - kl

        ctx.spec, ctx.s, ctx.ashape, ctx.gs, ctx.gw, ctx.gsaved, ctx.att = spec, s, ashape, gs, gw, gsaved, att
        ctx.spacing = spacing
        ctx.enc_shapes = [p.shape for p in enc_params]
        ctx.param_objs = (enc_params, gen_params) if spec.sync is not None else None   # for the fused optimiser epilogue
        ctx.save_for_backward(yc, w2m, wh, x1, h, heads, xc, gum, rz, rth, log_prior, y_hat, ctfc if ctfc is not None else yc,
                              *((enc_params[10], xp) if es.pool else ()))
        ctx.has_ctf = ctfc is not None
        return elbo, log_p, kl

    @staticmethod
    def backward(ctx, g_elbo, g_logp, g_kl):
        yc, w2m, wh, x1, h, heads, xc, gum, rz, rth, log_prior, y_hat, ctfc, *pool = ctx.saved_tensors
        spec, s, gs, att = ctx.spec, ctx.s, ctx.gs, ctx.att
        B = s.B
        zero = torch.zeros((), device=yc.device)
        g_elbo = zero if g_elbo is None else g_elbo
        g_logp = zero if g_logp is None else g_logp
        g_kl = zero if g_kl is None else g_kl
        # device scalars: dLoss/d(ll_b) and dLoss/d(kl_b) (no host sync)
        w_ll = ((g_elbo + g_logp) / B).reshape(1).float().contiguous()
        w_kl = ((g_kl - g_elbo) / B).reshape(1).float().contiguous()
        yflat = yc.reshape(B, -1)
        if spec.likelihood == "bernoulli":
            _, d_yhat = ops.bernoulli(y_hat, yflat, w_ll)
        elif spec.likelihood == "gaussian_fit_noise":
            _, d_yhat = ops.gaussian_fit_noise(y_hat, yflat, w_ll)
        else:
            _, d_yhat, _ = ops.gaussian(y_hat, yflat, s.n, ctfc if ctx.has_ctf else None, att["dx"], ctx.spacing,
                                        spec.mask_radius, w_ll, mu=ctx.mu)
            ctx.mu = None
        with nvtx_range("tvae.generator_bwd"):
            gout = ops.generator_bwd(gs, ctx.gw, xc, att["theta_b"], att["dx"], att["zb"], ctx.gsaved, y_hat, d_yhat)
        gen_grads = _gen_param_grads(gout, gs.L)
        if spec.sync is not None:
            spec.sync.start(0, gout["flat"], ctx.param_objs[1], gen_grads)
        with nvtx_range("tvae.attention_bwd"):
            d_heads = ops.attn_bwd(ctx.ashape, heads, gum, rz, rth, log_prior, att, gout["d_z"], gout["d_theta"], gout["d_dx"], w_kl)
        with nvtx_range("tvae.encoder_bwd"):
            enc_grads, enc_flat = _encoder_backward(s, spec.enc, yc, w2m, wh, x1, h, d_heads, ctx.enc_shapes, *pool)
        if spec.sync is not None:
            spec.sync.start(1, enc_flat, ctx.param_objs[0], enc_grads)
            spec.sync.finish()            # gen_grads / enc_grads are views of the two buckets: averaged in place
            ctx.param_objs = None
        ctx.gsaved = None
        ctx.att = None
        return (None,) * 9 + tuple(enc_grads) + tuple(gen_grads)
