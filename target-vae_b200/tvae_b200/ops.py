"""Thin torch-tensor wrappers over the C ABI (include/tvae_b200.h).  torch is plumbing only: it owns device
memory and the stream; every computation is a call into libtvae_b200.so."""
from __future__ import annotations

import ctypes
import functools
import math
from ctypes import POINTER, Structure, byref, c_float, c_int, c_void_p

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class EncShape(Structure):
    _fields_ = [(n, c_int) for n in ("B", "C", "n", "k", "p", "G", "O", "z", "kpad", "act")]


class EncFwdArgs(Structure):
    _fields_ = [(n, c_void_p) for n in ("y", "bank", "conv1_bias", "w2", "b2", "wh", "bh", "head_add",
                                         "x1", "h", "heads", "w2_h", "fc_w", "fc_b", "xp")]


class EncBwdArgs(Structure):
    _fields_ = [(n, c_void_p) for n in ("y", "w2", "wh", "x1", "h", "d_heads", "dhpre", "dx1_16", "w2t_h", "scales", "dbank",
                                         "dw2", "db2", "dwh", "dbh", "fc_w", "xp", "dxp16", "dfc_w", "dfc_b")]


class RefineArgs(Structure):
    _fields_ = ([(n, c_void_p) for n in ("y", "weight", "conv1_bias", "w2", "b2", "wh", "bh", "head_add", "fc_w", "fc_b", "heads")]
                + [("rel_tol", c_float)]
                + [(n, c_void_p) for n in ("bank32", "cand", "n_cand", "cand_heads", "z_content", "theta_mu", "argmax", "refined_logit")])


REFINE_MAX_CAND = 32          # TVAE_REFINE_MAX_CAND
REFINE_REL_TOL = 2e-3         # candidate band, as a fraction of the attention map's range


class AttnShape(Structure):
    _fields_ = [("B", c_int), ("G", c_int), ("d", c_int), ("z", c_int), ("s", c_float),
                ("theta_prior_std", c_float), ("offsets", c_float * 16)]


class AttnFwdArgs(Structure):
    _fields_ = [(n, c_void_p) for n in ("heads", "gumbel", "r_z", "r_theta", "log_prior", "stats", "zb",
                                         "theta_b", "dx", "kl")]


class AttnBwdArgs(Structure):
    _fields_ = [("f", AttnFwdArgs)] + [(n, c_void_p) for n in ("g_zb", "g_theta", "g_dx", "g_kl", "d_heads")]


class GenShape(Structure):
    _fields_ = [(n, c_int) for n in ("B", "N", "E", "H", "L", "n_out", "zdim", "act")]


class GenFwdArgs(Structure):
    _fields_ = [(n, c_void_p) for n in ("x", "theta", "dx", "z", "wf_scaled", "bf", "w1", "b1", "wz", "wh", "bh",
                                         "wout", "bout", "zb", "acts", "y_hat", "w_h", "mask_bits")]


class GenBwdArgs(Structure):
    _fields_ = [("f", GenFwdArgs)] + [(n, c_void_p) for n in (
        "d_yhat", "dpre0", "dpre1", "wt_h", "scales", "dxp", "dzb", "dw1", "db1", "dwz", "dwh", "dbh", "dwout", "dbout",
        "d_theta", "d_dx", "d_z")]


def _p(t):
    return None if t is None else t.data_ptr()


def _set(struct, **kw):
    """Fill pointer fields from tensors; the tensors are kept alive on the struct until it is dropped."""
    keep = struct.__dict__.setdefault("_keep", [])
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            keep.append(v)
        setattr(struct, k, _p(v) if (v is None or isinstance(v, torch.Tensor)) else v)
    return struct


_configured = False


def L():
    global _configured
    lib = _lib.lib()
    if not _configured:
        lib.tvae_bank_pitch.restype = c_int
        lib.tvae_bank16_pitch.restype = c_int
        lib.tvae_launch_count.restype = ctypes.c_longlong
        lib.tvae_profile_enable.restype = None
        lib.tvae_profile_collect.restype = c_int
        lib.tvae_conv1_executed_fraction.restype = ctypes.c_double
        lib.tvae_conv1_executed_fraction.argtypes = [c_void_p, c_int]
        for name in ("tvae_filter_bank_fwd", "tvae_filter_bank_bwd", "tvae_encoder_fwd", "tvae_encoder_bwd",
                     "tvae_attn_log_prior", "tvae_attn_fwd", "tvae_attn_bwd", "tvae_attn_softmax_pair",
                     "tvae_get_latent", "tvae_refine_argmax", "tvae_generator_fwd", "tvae_generator_bwd", "tvae_bernoulli",
                     "tvae_gaussian", "tvae_gaussian_fit_noise"):
            getattr(lib, name).restype = c_int
        lib.tvae_gaussian_fit_noise.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
        lib.tvae_gaussian.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_float, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
        lib.tvae_gaussian_workspace_bytes.restype = ctypes.c_longlong
        lib.tvae_gaussian_workspace_bytes.argtypes = [c_int, c_int]
        lib.tvae_bernoulli.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
        _configured = True
    return lib


def f32(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def empty(*shape, device):
    return torch.empty(*shape, device=device, dtype=torch.float32)


def half(*shape, device):
    return torch.empty(*shape, device=device, dtype=torch.float16)


def flat_views(shapes, device):
    """One flat fp32 buffer holding a tensor per shape (each starting on a 16-byte boundary) -> (flat, [views]).
    The backward kernels write the parameter gradients straight into such a bucket through the C ABI's output pointers,
    so that the data-parallel all-reduce (tvae_b200.dp.GradSync) runs on it in place - no flatten / concatenate copy."""
    sizes = [int(torch.Size(sh).numel()) for sh in shapes]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 3) // 4 * 4
    flat = torch.empty(max(total, 4), device=device, dtype=torch.float32)
    return flat, [flat[o:o + n].view(sh) for o, n, sh in zip(offs, sizes, shapes)]


# ----------------------------------------------------------------------------------------------- encoder
ACT_LEAKYRELU, ACT_TANH = 0, 1      # TVAE_ACT_* of include/tvae_b200.h


def act_kind(module) -> int:
    """nn.LeakyReLU() (default slope) / nn.Tanh() instance -> TVAE_ACT_*; anything else is not on the kernels."""
    if isinstance(module, torch.nn.Tanh):
        return ACT_TANH
    if isinstance(module, torch.nn.LeakyReLU) and module.negative_slope == 0.01:
        return ACT_LEAKYRELU
    raise NotImplementedError(f"activation {module!r}: the trainers' --activation offers leakyrelu and tanh only")


def enc_shape(B, C, n, k, p, G, O, z, act=ACT_LEAKYRELU) -> EncShape:
    return EncShape(B, C, n, k, p, G, O, z, L().tvae_bank_pitch(C, k), act)


def filter_bank_fwd(s: EncShape, weight: torch.Tensor) -> torch.Tensor:
    # fp16 [G*O][kpad16]: operand of the kind::f16 conv GEMM (same 11-bit significand as TF32)
    bank = torch.empty(s.G * s.O, L().tvae_bank16_pitch(s.C, s.k), device=weight.device, dtype=torch.float16)
    check(L().tvae_filter_bank_fwd(byref(s), ptr(f32(weight)), ptr(bank), stream_ptr()), "tvae_filter_bank_fwd")
    return bank


def filter_bank_bwd(s: EncShape, dbank: torch.Tensor, out=None):
    """out = (dweight (O,C,1,k,k), dbias (O)) views to write into (a gradient bucket), or None to allocate."""
    dw, db = out if out is not None else (empty(s.O, s.C, 1, s.k, s.k, device=dbank.device), empty(s.O, device=dbank.device))
    check(L().tvae_filter_bank_bwd(byref(s), ptr(dbank), ptr(dw), ptr(db), stream_ptr()), "tvae_filter_bank_bwd")
    return dw, db


@functools.lru_cache(maxsize=None)
def rotation_offsets(G: int, rot_refinement: bool):
    """models.py:361-366 / :401."""
    if not rot_refinement:
        return [0.0] * G
    out = []
    for i in range(G):
        a = i * 2 * math.pi / G
        if i > G // 2:
            a -= 2 * math.pi
        out.append(float(torch.tensor(a, dtype=torch.float32)))
    return out


@functools.lru_cache(maxsize=None)
def rotation_log_prior(G: int, rot_refinement: bool, normal_prior_over_r: bool, theta_prior: float):
    """p_r[r] of models.py:360-379 as python floats (fp32 arithmetic like the reference)."""
    if rot_refinement:
        offs = torch.tensor(rotation_offsets(G, True), dtype=torch.float32)
        if normal_prior_over_r:
            sigma = torch.tensor(theta_prior, dtype=torch.float32)
            p = -(offs ** 2) / (2 * sigma ** 2) - torch.log(sigma) - 0.5 * math.log(2 * math.pi)
        else:
            p = torch.zeros(G) - torch.log(torch.tensor(4 * math.pi, dtype=torch.float32))
    else:
        p = torch.zeros(G) - math.log(G)
    return [float(v) for v in p]


def head_tables(wa, ba, wr, br, wz, bz, G, p_r, offsets, device):
    """Stack the three 1x1x1 head convolutions and build the per-(channel, rotation) additive table."""
    O = wa.shape[1]
    wh = torch.cat([f32(wa).reshape(1, O), f32(wr).reshape(2, O), f32(wz).reshape(-1, O)], 0).contiguous()
    bh = torch.cat([f32(ba).reshape(1), f32(br).reshape(2), f32(bz).reshape(-1)], 0).contiguous()
    return wh, bh, _head_add_table(wh.shape[0], G, tuple(p_r), tuple(offsets), str(device))


@functools.lru_cache(maxsize=64)
def _head_add_table(NH, G, p_r, offsets, device):
    add = torch.zeros(NH, G, dtype=torch.float32)
    add[0] = torch.tensor(p_r, dtype=torch.float32)
    add[1] = torch.tensor(offsets, dtype=torch.float32)
    return add.to(device)


def encoder_fwd(s: EncShape, y, bank, b1, w2, b2, wh, bh, head_add, keep_h=True, pool=None):
    """keep_h = False: inference only (get_latent) - the hidden map h is neither allocated nor written.
    pool = (fc_r.weight, fc_r.bias): rotation pooling between conv1 and conv2 (attention/unimodal encoder with
    groupconv > 0, models.py:301-304); h and the head maps then have one rotation slot.  -> x1, h, heads, xp."""
    dev = y.device
    d = s.n + 2 * s.p - s.k + 1
    P = d * d
    R = s.B * s.G * P
    NH = 3 + 2 * s.z
    # activations are stored fp16 (the MMA operand format: 11-bit significand like TF32, half the HBM traffic)
    x1 = half(R, s.O, device=dev)
    G2 = 1 if pool is not None else s.G
    h = half(s.B * G2 * P, s.O, device=dev) if keep_h else None
    heads = empty(s.B, NH, G2, P, device=dev)
    w2r = half(s.O, s.O, device=dev)
    xp = half(s.B * P, s.O, device=dev) if pool is not None else None
    a = _set(EncFwdArgs(), y=f32(y), bank=bank, conv1_bias=f32(b1), w2=f32(w2), b2=f32(b2), wh=wh, bh=bh,
             head_add=head_add, x1=x1, h=h, heads=heads, w2_h=w2r,
             fc_w=None if pool is None else f32(pool[0]).reshape(-1), fc_b=None if pool is None else f32(pool[1]).reshape(-1), xp=xp)
    check(L().tvae_encoder_fwd(byref(s), byref(a), stream_ptr()), "tvae_encoder_fwd")
    return x1, h, heads, xp


def encoder_bwd(s: EncShape, y, w2, wh, x1, h, d_heads, pool=None, out=None):
    """pool = (fc_r.weight, xp) with rotation pooling: additionally returns (dfc_w (G), dfc_b (1)).
    out = (dw2 (O,O), db2 (O), dwh (NH,O), dbh (NH) [, dfc_w (G), dfc_b (1)]) views to write into, or None to allocate."""
    dev = y.device
    NH = 3 + 2 * s.z
    R = x1.shape[0]
    dhpre = half(h.shape[0], s.O, device=dev)
    w2t = half(s.O, s.O, device=dev)
    scales = empty(8, device=dev)
    dbank = empty(s.G * s.O, s.kpad, device=dev)
    if out is not None:
        dw2, db2, dwh, dbh = out[:4]
    else:
        dw2, db2, dwh, dbh = empty(s.O, s.O, device=dev), empty(s.O, device=dev), empty(NH, s.O, device=dev), empty(NH, device=dev)
    dx1_16 = half(R, s.O, device=dev)
    a = _set(EncBwdArgs(), y=f32(y), w2=f32(w2), wh=wh, x1=x1, h=h, d_heads=f32(d_heads), dhpre=dhpre, dx1_16=dx1_16, w2t_h=w2t,
             scales=scales,
             dbank=dbank, dw2=dw2, db2=db2, dwh=dwh, dbh=dbh)
    dfc_w = dfc_b = None
    if pool is not None:
        fc_w, xp = pool
        dfc_w, dfc_b = out[4:6] if out is not None else (empty(s.G, device=dev), empty(1, device=dev))
        _set(a, fc_w=f32(fc_w).reshape(-1), xp=xp, dxp16=half(xp.shape[0], s.O, device=dev), dfc_w=dfc_w, dfc_b=dfc_b)
    check(L().tvae_encoder_bwd(byref(s), byref(a), stream_ptr()), "tvae_encoder_bwd")
    if pool is not None:
        return dbank, dw2, db2, dwh, dbh, dfc_w, dfc_b
    return dbank, dw2, db2, dwh, dbh


# ----------------------------------------------------------------------------------------------- attention
def attn_shape(B, G, d, z, s, offsets, theta_prior_std=None) -> AttnShape:
    """theta_prior_std: None = pi / G (attention/attention branch, train_mnist.py:269-272)."""
    std = math.pi / G if theta_prior_std is None else theta_prior_std
    a = AttnShape(B, G, d, z, float(s), float(torch.tensor(std, dtype=torch.float32)))
    for i, o in enumerate(offsets):
        a.offsets[i] = o
    return a


_log_prior_cache: dict = {}


def attn_log_prior(s: AttnShape, p_r, device):
    """log p(t, r) depends only on the attention geometry and the rotation prior: computed once per configuration
    (the reference rebuilds it on the host every step, train_mnist.py:258-262)."""
    key = (s.G, s.d, float(s.s), tuple(float(v) for v in p_r), str(device))
    hit = _log_prior_cache.get(key)
    if hit is not None:
        return hit
    if len(_log_prior_cache) > 32:
        _log_prior_cache.clear()
    out = _attn_log_prior(s, p_r, device)
    _log_prior_cache[key] = out
    return out


def _attn_log_prior(s: AttnShape, p_r, device):
    out = empty(s.G * s.d * s.d, device=device)
    arr = (c_float * 16)(*([float(v) for v in p_r] + [0.0] * (16 - len(p_r))))
    check(L().tvae_attn_log_prior(byref(s), arr, ptr(out), stream_ptr()), "tvae_attn_log_prior")
    return out


def attn_fwd(s: AttnShape, heads, gumbel, r_z, r_theta, log_prior):
    dev = heads.device
    out = dict(stats=empty(s.B, 4, device=dev), zb=empty(s.B, s.z, device=dev), theta_b=empty(s.B, device=dev),
               dx=empty(s.B, 2, device=dev), kl=empty(s.B, device=dev))
    a = _set(AttnFwdArgs(), heads=heads, gumbel=f32(gumbel), r_z=f32(r_z), r_theta=f32(r_theta), log_prior=log_prior, **out)
    check(L().tvae_attn_fwd(byref(s), byref(a), stream_ptr()), "tvae_attn_fwd")
    return out


def attn_bwd(s: AttnShape, heads, gumbel, r_z, r_theta, log_prior, fwd_out, g_zb, g_theta, g_dx, g_kl):
    d_heads = torch.empty_like(heads)
    a = AttnBwdArgs()
    _set(a.f, heads=heads, gumbel=f32(gumbel), r_z=f32(r_z), r_theta=f32(r_theta), log_prior=log_prior, **fwd_out)
    _set(a, g_zb=f32(g_zb), g_theta=f32(g_theta), g_dx=f32(g_dx), g_kl=f32(g_kl), d_heads=d_heads)
    check(L().tvae_attn_bwd(byref(s), byref(a), stream_ptr()), "tvae_attn_bwd")
    return d_heads


def attn_softmax_pair(heads, gumbel):
    B, NH = heads.shape[0], heads.shape[1]
    Lr = heads.shape[2] * heads.shape[3]
    q = empty(B, Lr, device=heads.device)
    a = empty(B, Lr, device=heads.device)
    check(L().tvae_attn_softmax_pair(ptr(heads), ptr(f32(gumbel)), ptr(q), ptr(a), B, NH, Lr, stream_ptr()),
          "tvae_attn_softmax_pair")
    return q, a


def get_latent(s: AttnShape, heads):
    dev = heads.device
    zc = empty(s.B, 2 * s.z, device=dev)
    th = empty(s.B, 1, device=dev)
    dx = empty(s.B, 2, device=dev)
    am = torch.empty(s.B, device=dev, dtype=torch.int32)
    check(L().tvae_get_latent(byref(s), ptr(heads), ptr(zc), ptr(th), ptr(dx), ptr(am), stream_ptr()), "tvae_get_latent")
    return zc, th, dx, am


def refine_argmax(s: EncShape, y, weight, b1, w2, b2, wh, bh, head_add, heads, pool=None, rel_tol=REFINE_REL_TOL):
    """Exact re-evaluation of the logit chain at the near-maximal cells of the fast attention map (tvae_refine_argmax).
    heads: fast maps (B, NH, G2, P).  -> dict(argmax (B) int32, z_content (B,2z), theta_mu (B,1), n_cand (B) int32,
    cand (B,32) int32, cand_heads (B,32,NH), logit (B))."""
    dev = heads.device
    B, NH = heads.shape[0], heads.shape[1]
    K = s.C * s.k * s.k
    out = dict(argmax=torch.empty(B, device=dev, dtype=torch.int32), z_content=empty(B, 2 * s.z, device=dev),
               theta_mu=empty(B, 1, device=dev), n_cand=torch.empty(B, device=dev, dtype=torch.int32),
               cand=torch.zeros(B, REFINE_MAX_CAND, device=dev, dtype=torch.int32),
               cand_heads=torch.zeros(B, REFINE_MAX_CAND, NH, device=dev, dtype=torch.float32), logit=empty(B, device=dev))
    bank32 = empty(s.G * s.O, K, device=dev)
    a = _set(RefineArgs(), y=f32(y), weight=f32(weight), conv1_bias=f32(b1), w2=f32(w2), b2=f32(b2), wh=wh, bh=bh, head_add=head_add,
             fc_w=None if pool is None else f32(pool[0]).reshape(-1), fc_b=None if pool is None else f32(pool[1]).reshape(-1),
             heads=heads, bank32=bank32, cand=out["cand"], n_cand=out["n_cand"], cand_heads=out["cand_heads"],
             z_content=out["z_content"], theta_mu=out["theta_mu"], argmax=out["argmax"], refined_logit=out["logit"])
    a.rel_tol = float(rel_tol)
    check(L().tvae_refine_argmax(byref(s), byref(a), stream_ptr()), "tvae_refine_argmax")
    return out


# ----------------------------------------------------------------------------------------------- generator
class GenWeights:
    """Flat views of SpatialGenerator parameters in the order the C ABI expects."""

    def __init__(self, wf_scaled, bf, w1, b1, wz, hidden_w, hidden_b, wout, bout):
        self.wf_scaled, self.bf, self.w1, self.b1, self.wz = wf_scaled, bf, f32(w1), f32(b1), f32(wz)
        self.L = len(hidden_w)
        H = self.w1.shape[0]
        dev = self.w1.device
        self.wh = torch.stack([f32(w) for w in hidden_w]).contiguous() if self.L else torch.zeros(1, device=dev)
        self.bh = torch.stack([f32(b) for b in hidden_b]).contiguous() if self.L else torch.zeros(1, device=dev)
        self.wout, self.bout = f32(wout), f32(bout)
        self.H = H
        self.E = 0 if wf_scaled is None else wf_scaled.shape[0]


def gen_shape(B, N, gw: GenWeights, zdim, act=ACT_LEAKYRELU) -> GenShape:
    return GenShape(B, N, gw.E, gw.H, gw.L, gw.wout.shape[0], zdim, act)


def _gen_fwd_args(s: GenShape, gw: GenWeights, x, theta, dx, z, zb, acts, y_hat, w_h, mask_bits=None):
    return _set(GenFwdArgs(), x=f32(x), theta=None if theta is None else f32(theta), dx=None if dx is None else f32(dx),
                z=f32(z), wf_scaled=gw.wf_scaled, bf=gw.bf, w1=gw.w1, b1=gw.b1, wz=gw.wz, wh=gw.wh, bh=gw.bh,
                wout=gw.wout, bout=gw.bout, zb=zb, acts=acts, y_hat=y_hat, w_h=w_h, mask_bits=mask_bits)


def generator_fwd(s: GenShape, gw: GenWeights, x, theta, dx, z):
    dev = z.device
    M = s.B * s.N
    zb = empty(s.B, s.H, device=dev)
    acts = half(s.L + 1, M, s.H, device=dev)        # fp16 activations: the MMA operand format, half the HBM traffic
    y_hat = empty(M, s.n_out, device=dev)
    w_h = half(s.H * max(s.E, 2) + s.L * s.H * s.H, device=dev)
    # one-bit LeakyReLU masks of acts[0 .. L-1] for the hidden layers' input gradients (1/16 of the activation's bytes)
    mask_bits = None
    if s.act == ACT_LEAKYRELU and s.H % 64 == 0 and s.L >= 1:
        mask_bits = torch.empty(s.L * (s.H // 64) * M, device=dev, dtype=torch.int64)
    a = _gen_fwd_args(s, gw, x, theta, dx, z, zb, acts, y_hat, w_h, mask_bits)
    check(L().tvae_generator_fwd(byref(s), byref(a), stream_ptr()), "tvae_generator_fwd")
    return y_hat, dict(zb=zb, acts=acts, w_h=w_h, mask_bits=mask_bits)


def generator_bwd(s: GenShape, gw: GenWeights, x, theta, dx, z, saved, y_hat, d_yhat):
    dev = z.device
    M = s.B * s.N
    H, E, Lh = s.H, s.E, s.L
    # every parameter gradient lives in ONE flat bucket (out["flat"]): the data-parallel all-reduce runs on it in place
    names = ("dw1", "db1", "dwz", "dwh", "dbh", "dwout", "dbout")
    flat, views = flat_views([(H, max(E, 2)), (H,), (H, s.zdim), (max(Lh, 1), H, H), (max(Lh, 1), H), (s.n_out, H), (s.n_out,)], dev)
    out = dict(zip(names, views))
    out.update(flat=flat, d_theta=empty(s.B, device=dev), d_dx=empty(s.B, 2, device=dev), d_z=empty(s.B, s.zdim, device=dev))
    scratch = dict(dpre0=half(M, H, device=dev), dpre1=half(M, H, device=dev), wt_h=half(max(E * H, H * H), device=dev),
                   scales=empty(32, device=dev), dxp=empty(M, 2, device=dev), dzb=empty(s.B, H, device=dev))
    a = GenBwdArgs()
    a.f = _gen_fwd_args(s, gw, x, theta, dx, z, saved["zb"], saved["acts"], y_hat, saved["w_h"], saved.get("mask_bits"))
    _set(a, d_yhat=f32(d_yhat), **scratch, **{k: v for k, v in out.items() if k != "flat"})
    if theta is None:
        a.d_theta = None
        a.d_dx = None
    check(L().tvae_generator_bwd(byref(s), byref(a), stream_ptr()), "tvae_generator_bwd")
    out["dxp"] = scratch["dxp"]
    out["scales"] = scratch["scales"]      # the per-layer power-of-two scales chosen on the device (diagnostics)
    return out


# ----------------------------------------------------------------------------------------------- likelihoods
def bernoulli(y_hat, y, g=None):
    B = y.shape[0]
    E = y.numel() // B
    ll = empty(B, device=y.device)
    d = torch.empty_like(y_hat) if g is not None else None
    check(L().tvae_bernoulli(_p(f32(y_hat)), _p(f32(y)), _p(ll), _p(d), _p(g), B, E, stream_ptr().value), "tvae_bernoulli")
    return ll, d


def ctf_filter_size(ctf, B, n) -> int:
    """Validates a CTF filter stack against the minibatch and returns its size m.  The reference applies
    F.conv2d(y_mu.view(1,B,n,n), ctf, padding=ctf.size(2)//2, groups=B) (train_particles.py:298-302): one square filter
    per image, any odd size (m = n - 1 by default; with --crop the filters keep the uncropped size)."""
    if ctf.dim() == 4 and ctf.shape[1] == 1:
        ctf = ctf[:, 0]
    if ctf.dim() != 3 or ctf.shape[0] != B or ctf.shape[1] != ctf.shape[2]:
        raise ValueError(f"ctf must be (B,1,m,m) or (B,m,m) with B = {B}, got {tuple(ctf.shape)}")
    m = int(ctf.shape[-1])
    if m != n - 1 and m % 2 == 0:
        raise ValueError(f"ctf filter size {m} must be odd (or image_dim - 1 = {n - 1}): with an even size the reference's "
                         "grouped convolution changes the image size")
    return m


def gaussian(y_hat, y, n, ctf=None, dx=None, s=1.0, radius=0, g=None, mu=None, use_ctf_gemm=True):
    """-> (ll, d_yhat, mu).  `mu` = ctf (*) y_hat of an earlier call lets the backward pass skip the forward CTF."""
    B = y.shape[0]
    dev = y.device
    ll = empty(B, device=dev)
    m = 0 if ctf is None else ctf_filter_size(ctf, B, n)
    have_mu = mu is not None and ctf is not None
    if ctf is not None and mu is None:
        mu = empty(B, n * n, device=dev)
    dmu = empty(B, n * n, device=dev) if (ctf is not None and g is not None) else None
    d = torch.empty_like(y_hat) if g is not None else None
    ws = None
    if ctf is not None and use_ctf_gemm and m == n - 1:
        nbytes = int(L().tvae_gaussian_workspace_bytes(B, n))
        if nbytes > 0:
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    check(L().tvae_gaussian(_p(None if have_mu else f32(y_hat)), _p(f32(y)), _p(None if ctf is None else f32(ctf)), m,
                            _p(None if dx is None else f32(dx)), float(s), int(radius), _p(mu), _p(dmu), _p(ll), _p(d),
                            _p(g), B, n, _p(ws), stream_ptr().value), "tvae_gaussian")
    return ll, d, mu


def gaussian_fit_noise(y_hat2, y, g=None):
    """--fit-noise likelihood (train_particles.py:289-296, 333-334): y_hat2 (B*N, 2) generator output, y (B, N).
    -> (ll (B), d_yhat2 or None).  The reference's flat-halves reading of the (B, N, 2) output is kept."""
    B, N = y.shape[0], y.shape[1]
    if y_hat2.numel() != 2 * B * N:
        raise ValueError(f"gaussian_fit_noise: generator output has {y_hat2.numel()} values, expected {2 * B * N}")
    ll = empty(B, device=y.device)
    d = torch.empty_like(y_hat2) if g is not None else None
    check(L().tvae_gaussian_fit_noise(_p(f32(y_hat2)), _p(f32(y)), _p(ll), _p(d), _p(g), B, N, stream_ptr().value),
          "tvae_gaussian_fit_noise")
    return ll, d


# ----------------------------------------------------------------------------------------------- standalone module interfaces
def _module_lib():
    lib = L()
    if not getattr(lib, "_tvae_module_configured", False):
        ll, i, vp = ctypes.c_longlong, c_int, c_void_p
        lib.tvae_fourier_embed_fwd.restype = c_int
        lib.tvae_fourier_embed_fwd.argtypes = [vp, vp, vp, vp, ll, i, vp]
        lib.tvae_fourier_embed_bwd.restype = c_int
        lib.tvae_fourier_embed_bwd.argtypes = [vp, vp, vp, vp, vp, ll, i, vp]
        lib.tvae_linear_act_fwd.restype = c_int
        lib.tvae_linear_act_fwd.argtypes = [vp, vp, vp, i, i, i, i, i, vp, vp, vp, vp]
        lib.tvae_linear_act_bwd.restype = c_int
        lib.tvae_linear_act_bwd.argtypes = [vp, vp, vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp]
        lib.tvae_attn_softmax_pair_bwd.restype = c_int
        lib.tvae_attn_softmax_pair_bwd.argtypes = [vp, vp, vp, vp, vp, i, i, vp]
        lib._tvae_module_configured = True
    return lib


def fourier_embed_fwd(x, w_scaled, b):
    """RandomFourierEmbedding2d.forward: x (M,2) -> (M,E)."""
    M, E = x.shape[0], w_scaled.shape[0]
    out = empty(M, E, device=x.device)
    check(_module_lib().tvae_fourier_embed_fwd(_p(x), _p(w_scaled), _p(b), _p(out), M, E, stream_ptr().value), "tvae_fourier_embed_fwd")
    return out


def fourier_embed_bwd(x, w_scaled, b, g):
    M, E = x.shape[0], w_scaled.shape[0]
    dx = empty(M, 2, device=x.device)
    check(_module_lib().tvae_fourier_embed_bwd(_p(x), _p(w_scaled), _p(b), _p(g), _p(dx), M, E, stream_ptr().value), "tvae_fourier_embed_bwd")
    return dx


def linear_act_fwd(x, w, bias, resid, act):
    """act(x W^T + b [+ x]) on the tensor-core LinearNT kernel -> (y (M,N) fp32, x16 saved for the backward)."""
    M, K = x.shape
    N = w.shape[0]
    y = empty(M, N, device=x.device)
    x16, w16 = half(M, K, device=x.device), half(N, K, device=x.device)
    check(_module_lib().tvae_linear_act_fwd(_p(x), _p(w), _p(bias), M, N, K, int(resid), int(act), _p(y), _p(x16), _p(w16),
                                            stream_ptr().value), "tvae_linear_act_fwd")
    return y, x16


def linear_act_bwd(x16, w, y, g, resid, act, need_dx=True):
    M, K = x16.shape
    N = w.shape[0]
    dev = w.device
    dx = empty(M, K, device=dev) if need_dx else None
    dw, db = empty(N, K, device=dev), empty(N, device=dev)
    dpre16, wt16, scales = half(M, N, device=dev), half(K, N, device=dev), empty(8, device=dev)
    check(_module_lib().tvae_linear_act_bwd(_p(x16), _p(w), _p(y), _p(g), M, N, K, int(resid), int(act), _p(dpre16), _p(wt16), _p(scales),
                                            _p(dx), _p(dw), _p(db), stream_ptr().value), "tvae_linear_act_bwd")
    return dx, dw, db


def attn_softmax_pair_bwd(q, a, dq, da):
    B, Lr = q.shape
    d = empty(B, Lr, device=q.device)
    check(_module_lib().tvae_attn_softmax_pair_bwd(_p(q), _p(a), _p(dq), _p(da), _p(d), B, Lr, stream_ptr().value), "tvae_attn_softmax_pair_bwd")
    return d


# ----------------------------------------------------------------------------------------------- instrumentation
def launch_count() -> int:
    return int(L().tvae_launch_count())


# kernels (names of tvae_profile_collect) whose MMAs run with fp16 operands (kind::f16): all of them
F16_KERNELS: set = {"conv1_fwd", "conv1_wgrad", "conv2_heads", "gen_l1_fwd", "gen_l1_wgrad", "gen_l1_dgrad", "linear_nt", "linear_tn"}


def conv1_executed_fraction(s: EncShape, wgrad: bool) -> float:
    """Executed / dense-count K chunks of the conv1 kernels (zero-padding chunks are skipped); host arithmetic."""
    return float(L().tvae_conv1_executed_fraction(ctypes.cast(ctypes.pointer(s), c_void_p), 1 if wgrad else 0))


_profiling = False


def profile_enable(on: bool) -> None:
    global _profiling
    _profiling = bool(on)
    L().tvae_profile_enable(1 if on else 0)


def profile_enabled() -> bool:
    return _profiling


def profile_collect():
    """{kernel name: (total ms, launches)} of the tensor-core GEMM launches since profile_enable(True)."""
    cap = 32
    names = (ctypes.c_char_p * cap)()
    ms = (c_float * cap)()
    cnt = (c_int * cap)()
    n = L().tvae_profile_collect(names, ms, cnt, cap)
    return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}
