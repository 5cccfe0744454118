"""CPU oracle for the TARGET-VAE training hot path.  TEST INFRASTRUCTURE ONLY.

This file is a dtype-generic (fp32 / fp64) restatement, in plain PyTorch CPU ops,
of the reference algorithm for the path named in BASELINE.json `north_star`:
P_G group-conv encoder -> attention (t, r) inference (+offsets) -> ELBO terms ->
coordinate-MLP generator -> likelihood.  Gradients come from torch autograd over
this restatement.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  The product path
(`target-vae_b200/`) never does: it calls the sm_100a kernels through the C ABI
and fails loudly without them.

Parity pin: the reference ships no tests / golden vectors, so
the oracle is pinned against outputs of the *reference itself* executed in the
build container: `oracle/make_golden.py` imports `/root/reference` unmodified,
runs it on seeded inputs and commits the results under `tests/golden/`;
`tests/test_oracle_golden.py` checks this file against those fixtures.

Citations are `file:line` relative to the reference checkout.

Known, documented deviations (all far below the parity tolerance):
  * the translation grid is built in the working dtype as (j - d//2) * s rather
    than through float64 `np.arange` (train_mnist.py:209-218); |diff| ~ 1e-8.
  * the rotated filter bank is evaluated from the closed-form bilinear formula,
    not through `affine_grid`/`grid_sample`; |diff| <= 2.3e-7 in fp32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.01  # nn.LeakyReLU default, models.py:66,330
EPS_STD = 1e-6      # train_mnist.py:197


# ----------------------------------------------------------------------------
# a-1  GroupConv.trans_filter (models.py:174-197)
# ----------------------------------------------------------------------------
def rotation_angles(G: int):
    """theta_i accumulated in Python double: theta += 2*pi/G (models.py:181-195)."""
    d_theta = 2 * math.pi / G
    out, theta = [], 0.0
    for _ in range(G):
        out.append(theta)
        theta += d_theta
    return out


def rotated_filter_bank(weight: torch.Tensor, G: int) -> torch.Tensor:
    """weight (O,C,1,k,k) -> bank (O,G,C,1,k,k).

    bank[o,i,c,0,v,u] = bilinear sample (zero padding, align_corners=False) of
    weight[o,c,0] at ix = cos*cx + sin*cy + c0, iy = -sin*cx + cos*cy + c0 with
    cx = u - c0, cy = v - c0, c0 = (k-1)/2  (models.py:186-193 through
    affine_grid/grid_sample semantics).  cos/sin are rounded to fp32 first, as the
    reference stores them in an fp32 `rot` tensor (models.py:186-190).
    """
    O, C, D, k, k2 = weight.shape
    assert D == 1 and k == k2
    dt = weight.dtype
    c0 = (k - 1) / 2.0
    u = torch.arange(k, dtype=dt, device=weight.device) - c0
    cy, cx = torch.meshgrid(u, u, indexing="ij")  # cy varies with v (rows)
    w2 = weight[:, :, 0]  # (O,C,k,k)
    banks = []
    for theta in rotation_angles(G):
        c = float(torch.tensor(math.cos(theta), dtype=torch.float32))
        s = float(torch.tensor(math.sin(theta), dtype=torch.float32))
        ix = c * cx + s * cy + c0
        iy = -s * cx + c * cy + c0
        x0 = torch.floor(ix)
        y0 = torch.floor(iy)
        fx = ix - x0
        fy = iy - y0
        acc = torch.zeros(O, C, k, k, dtype=dt, device=weight.device)
        for dy, wy in ((0, 1 - fy), (1, fy)):
            for dx, wx in ((0, 1 - fx), (1, fx)):
                xi = (x0 + dx).long()
                yi = (y0 + dy).long()
                ok = (xi >= 0) & (xi < k) & (yi >= 0) & (yi < k)
                xi_c = xi.clamp(0, k - 1)
                yi_c = yi.clamp(0, k - 1)
                vals = w2[:, :, yi_c, xi_c]  # (O,C,k,k)
                acc = acc + vals * (wy * wx * ok.to(dt))
        banks.append(acc)
    bank = torch.stack(banks, dim=1)  # (O,G,C,k,k)
    return bank.unsqueeze(3)


# ----------------------------------------------------------------------------
# a-2  GroupConv.forward (models.py:202-225)
# ----------------------------------------------------------------------------
def groupconv_forward(y, weight, bias, G: int, padding: int):
    """y (B,C,n,n) -> (B,O,G,H',W'); flattened conv channel index = o*G + r."""
    O, C, _, k, _ = weight.shape
    tw = rotated_filter_bank(weight, G).reshape(O * G, C, k, k)
    out = F.conv2d(y.reshape(y.shape[0], C, y.shape[-2], y.shape[-1]), tw, None, 1, padding)
    B, _, Ho, Wo = out.shape
    out = out.view(B, O, G, Ho, Wo)
    if bias is not None:
        out = out + bias.view(1, O, 1, 1, 1)
    return out


# ----------------------------------------------------------------------------
# a-3  InferenceNetwork_AttentionTranslation_AttentionRotation.forward
#      (models.py:354-403)
# ----------------------------------------------------------------------------
def rotation_offsets(G: int, rot_refinement: bool, dtype=torch.float32):
    """models.py:361-366 (values wrapped to (-pi, pi]) / :401 (zeros)."""
    if not rot_refinement:
        return torch.zeros(G, dtype=dtype)
    vals = []
    for i in range(G):
        a = i * 2 * math.pi / G
        # reference literal lists: 0, pi/G*2 ... pi, then negative angles
        if i > G // 2:
            a = a - 2 * math.pi
        vals.append(a)
    # the reference builds an fp32 tensor (`.type(torch.float)`)
    return torch.tensor(vals, dtype=torch.float32).to(dtype)


def rotation_log_prior(G: int, rot_refinement: bool, normal_prior_over_r: bool,
                       theta_prior: float, dtype=torch.float32):
    """p_r (G,1,1), models.py:360-379."""
    if rot_refinement:
        offs = rotation_offsets(G, True, dtype)
        if normal_prior_over_r:
            sigma = torch.tensor(theta_prior, dtype=torch.float32).to(dtype)
            p = -(offs ** 2) / (2 * sigma ** 2) - torch.log(sigma) - 0.5 * math.log(2 * math.pi)
        else:
            lo = torch.tensor(-2 * math.pi, dtype=torch.float32).to(dtype)
            hi = torch.tensor(2 * math.pi, dtype=torch.float32).to(dtype)
            p = torch.zeros(G, dtype=dtype) - torch.log(hi - lo)
    else:
        p = torch.zeros(G, dtype=dtype) - math.log(G)
    return p.view(G, 1, 1)


@dataclass
class EncoderParams:
    conv1_w: torch.Tensor  # (O,C,1,k,k)
    conv1_b: torch.Tensor  # (O,)
    conv2_w: torch.Tensor  # (O,O,1,1,1)
    conv2_b: torch.Tensor
    conv_a_w: torch.Tensor  # (1,O,1,1,1)
    conv_a_b: torch.Tensor
    conv_r_w: torch.Tensor  # (2,O,1,1,1)
    conv_r_b: torch.Tensor
    conv_z_w: torch.Tensor  # (2z,O,1,1,1)
    conv_z_b: torch.Tensor
    activation: str = "leakyrelu"   # --activation (train_mnist.py:516-519): leakyrelu | tanh
    fc_r_w: Optional[torch.Tensor] = None   # (1,G)  rotation pooling of the attention/unimodal encoder, models.py:284
    fc_r_b: Optional[torch.Tensor] = None   # (1,)

    @staticmethod
    def from_state_dict(sd, dtype=torch.float32, activation="leakyrelu"):
        g = lambda k: sd[k].detach().clone().to(dtype)
        return EncoderParams(g("conv1.weight"), g("conv1.bias"), g("conv2.weight"), g("conv2.bias"),
                             g("conv_a.weight"), g("conv_a.bias"), g("conv_r.weight"), g("conv_r.bias"),
                             g("conv_z.weight"), g("conv_z.bias"), activation,
                             g("fc_r.weight") if "fc_r.weight" in sd else None,
                             g("fc_r.bias") if "fc_r.bias" in sd else None)

    def tensors(self):
        t = [self.conv1_w, self.conv1_b, self.conv2_w, self.conv2_b, self.conv_a_w, self.conv_a_b,
             self.conv_r_w, self.conv_r_b, self.conv_z_w, self.conv_z_b]
        return t + ([self.fc_r_w, self.fc_r_b] if self.fc_r_w is not None else [])

    # state_dict names in tensors() order (fc_r.* only when present: zip() with tensors() stops at the shorter list)
    names = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv_a.weight", "conv_a.bias",
             "conv_r.weight", "conv_r.bias", "conv_z.weight", "conv_z.bias", "fc_r.weight", "fc_r.bias"]


def _conv1x1(x, w, b):
    """nn.Conv3d(kernel 1) as a channel contraction; x (B,Cin,G,H,W)."""
    return torch.einsum("bcghw,oc->boghw", x, w.reshape(w.shape[0], w.shape[1])) + b.view(1, -1, 1, 1, 1)


def encoder_head_maps(y, p: EncoderParams, G: int, padding: int):
    """Returns raw head maps before priors/offsets: attn_raw (B,G,H',W'),
    theta_raw (B,2,G,H',W'), z (B,2z,G,H',W').  models.py:355-358,390-392."""
    x = _act(groupconv_forward(y, p.conv1_w, p.conv1_b, G, padding), p.activation)
    h = _act(_conv1x1(x, p.conv2_w, p.conv2_b), p.activation)
    attn = _conv1x1(h, p.conv_a_w, p.conv_a_b).squeeze(1)
    theta = _conv1x1(h, p.conv_r_w, p.conv_r_b)
    z = _conv1x1(h, p.conv_z_w, p.conv_z_b)
    return attn, theta, z


def encoder_forward(y, p: EncoderParams, G: int, padding: int, rot_refinement: bool,
                    normal_prior_over_r: bool, theta_prior: float, gumbel: torch.Tensor):
    """Full 7-tuple of models.py:403.  `gumbel` (B, G*H'*W') is the Gumbel noise
    -log(Exp(1)) that F.gumbel_softmax (tau=1, soft) draws at models.py:387."""
    dt = y.dtype
    attn, theta, z = encoder_head_maps(y, p, G, padding)
    p_r = rotation_log_prior(G, rot_refinement, normal_prior_over_r, theta_prior, dt)
    attn = attn + p_r
    B = attn.shape[0]
    q_t_r = F.log_softmax(attn.reshape(B, -1), dim=1).view_as(attn)
    a_sampled = F.softmax(attn.reshape(B, -1) + gumbel.reshape(B, -1), dim=1).view_as(attn)
    offsets = rotation_offsets(G, rot_refinement, dt)
    if rot_refinement:
        theta_mu = theta[:, 0] + offsets.view(1, G, 1, 1)
        theta = torch.stack((theta_mu, theta[:, 1]), dim=1)
    return attn, q_t_r, p_r, a_sampled, offsets, theta, z


# ----------------------------------------------------------------------------
# f-4  InferenceNetwork_AttentionTranslation_UnimodalRotation, groupconv == 0 (models.py:281-287, 300-318):
#      plain Conv2d(C, O, n, padding = n // 2) -> LeakyReLU -> 1x1 Conv2d -> LeakyReLU -> 1x1 heads.
#      Weights keep the module's 4-D Conv2d shapes; the maps carry a rotation axis of size 1 so that the
#      attention/attention code below (R = 1, no rotation prior, no offsets) restates
#      train_mnist.py:88-183 as well.
# ----------------------------------------------------------------------------
def plainconv_head_maps(y, p: EncoderParams, padding: int):
    if p.fc_r_w is not None:
        # groupconv = G > 0 (models.py:281-285, 301-304): P_G group conv, activation, then nn.Linear(G, 1) over the
        # rotation axis - (B,O,G,H',W') -> (B,O,H',W'); conv2 follows without an activation in between
        G = p.fc_r_w.shape[1]
        x = _act(groupconv_forward(y, p.conv1_w, p.conv1_b, G, padding), p.activation)
        x = (F.linear(x.permute(0, 1, 3, 4, 2), p.fc_r_w, p.fc_r_b).squeeze(4)).unsqueeze(2)
    else:
        x = _act(F.conv2d(y, p.conv1_w, p.conv1_b, padding=padding), p.activation).unsqueeze(2)
    h = _act(_conv1x1(x, p.conv2_w, p.conv2_b), p.activation)
    attn = _conv1x1(h, p.conv_a_w, p.conv_a_b).squeeze(1)     # (B,1,H',W')
    theta = _conv1x1(h, p.conv_r_w, p.conv_r_b)               # (B,2,1,H',W')
    z = _conv1x1(h, p.conv_z_w, p.conv_z_b)                   # (B,2z,1,H',W')
    return attn, theta, z


def plainconv_encoder_forward(y, p: EncoderParams, padding: int, gumbel: torch.Tensor):
    """models.py:300-318 in the 7-tuple layout of encoder_forward: q_t = log_softmax(attn) is what
    train_mnist.py:148 computes from the module's `attn`; p_r = 0 and offsets = 0 (one rotation, no prior)."""
    dt = y.dtype
    attn, theta, z = plainconv_head_maps(y, p, padding)
    B = attn.shape[0]
    q_t = F.log_softmax(attn.reshape(B, -1), dim=1).view_as(attn)
    a_sampled = F.softmax(attn.reshape(B, -1) + gumbel.reshape(B, -1), dim=1).view_as(attn)
    return attn, q_t, torch.zeros(1, 1, 1, dtype=dt), a_sampled, torch.zeros(1, dtype=dt), theta, z


# ----------------------------------------------------------------------------
# a-4  eval_minibatch, attention/attention(+offsets) branch
#      (train_mnist.py:187-282, train_particles.py:186-280)
# ----------------------------------------------------------------------------
def translation_grid(d: int, s, dtype=torch.float32):
    """(d*d, 2) grid of train_mnist.py:209-217: cell (i,j) -> ((j - d//2) s, (d-1-i - d//2) s)."""
    j = torch.arange(d, dtype=dtype)
    gx = (j - (d // 2)) * s
    gy = gx.flip(0)
    x1, x0 = torch.meshgrid(gy, gx, indexing="ij")  # x0[i,j] = gx[j], x1[i,j] = gy[i]
    return torch.stack([x0.reshape(-1), x1.reshape(-1)], dim=1)


def _kl_normal(mu_q, std_q, mu_p, std_p):
    """torch.distributions.kl._kl_normal_normal."""
    var_ratio = (std_q / std_p) ** 2
    t1 = ((mu_q - mu_p) / std_p) ** 2
    return 0.5 * (var_ratio + t1 - 1 - torch.log(var_ratio))


def attention_posterior(enc_out, x_coord, r_z, r_theta, G: int, groupconv_flag: int, theta_prior: float):
    """From the encoder 7-tuple to (z_b, theta_b, dx_b, x_transformed, kl_div).

    r_z (B,z,1) and r_theta (B,1,1) are the N(0,1) draws of train_mnist.py:206,230.
    """
    attn, q_t_r, p_r, a_sampled, offsets, theta_vals, z_vals = enc_out
    dt = attn.dtype
    B, R, d, _ = attn.shape
    s = (x_coord[1, 0] - x_coord[0, 0]).to(dt)
    x = x_coord.to(dt).expand(B, x_coord.shape[0], 2)

    a_locs = a_sampled.sum(1).reshape(B, -1, 1)
    a = a_sampled.reshape(B, -1).unsqueeze(2)
    z_vals = z_vals.reshape(B, z_vals.shape[1], -1)
    theta_vals = theta_vals.reshape(B, 2, -1)
    zd = z_vals.shape[1] // 2
    z_mu, z_logstd = z_vals[:, :zd], z_vals[:, zd:]
    z_std = torch.exp(z_logstd) + EPS_STD
    z_b = (torch.bmm(z_std, a) * r_z + torch.bmm(z_mu, a)).squeeze(2)

    grid = translation_grid(d, s, dt)  # (d*d,2)
    dx = torch.bmm(grid.t().unsqueeze(0).expand(B, 2, d * d), a_locs).squeeze(2).unsqueeze(1)  # (B,1,2)
    x = x - dx

    th_mu, th_logstd = theta_vals[:, 0:1], theta_vals[:, 1:2]
    th_std = torch.exp(th_logstd) + EPS_STD
    theta_b = (torch.bmm(th_std, a) * r_theta + torch.bmm(th_mu, a)).squeeze(2).squeeze(1)

    c, sn = torch.cos(theta_b), torch.sin(theta_b)
    rot = torch.stack([torch.stack([c, sn], 1), torch.stack([-sn, c], 1)], 1)  # (B,2,2)
    x = torch.bmm(x, rot)

    # KL terms (train_mnist.py:242-282)
    eq = torch.exp(q_t_r)
    dead = (eq == 0)
    z_mu5 = z_mu.reshape(B, zd, R, d, d)
    z_std5 = z_std.reshape(B, zd, R, d, d)
    z_mu5 = torch.where(dead.unsqueeze(1), torch.zeros_like(z_mu5), z_mu5)
    z_std5 = torch.where(dead.unsqueeze(1), torch.ones_like(z_std5), z_std5)
    th_mu4 = torch.where(dead, torch.zeros_like(q_t_r), th_mu.reshape(B, R, d, d))
    th_std4 = torch.where(dead, torch.ones_like(q_t_r), th_std.reshape(B, R, d, d))

    sig_t = torch.tensor(0.1, dtype=torch.float32).to(dt)
    log_p_t = (-(grid ** 2) / (2 * sig_t ** 2) - torch.log(sig_t) - 0.5 * math.log(2 * math.pi)).sum(1)
    p_t_r = log_p_t.view(1, 1, d, d) + p_r.unsqueeze(0)
    p_t_r = F.log_softmax(p_t_r.reshape(-1), dim=0).view(1, R, d, d)
    val1 = (eq * (q_t_r - p_t_r)).reshape(B, -1).sum(1)

    kl_z = _kl_normal(z_mu5, z_std5, torch.zeros((), dtype=dt), torch.ones((), dtype=dt)).sum(1)
    if groupconv_flag >= 1:
        th_prior_r = torch.tensor(math.pi / groupconv_flag, dtype=torch.float32).to(dt)
    else:
        th_prior_r = torch.tensor(theta_prior, dtype=torch.float32).to(dt)
    kl_t = _kl_normal(th_mu4, th_std4, offsets.to(dt).view(1, R, 1, 1), th_prior_r)
    val2 = (eq * (kl_t + kl_z)).reshape(B, -1).sum(1)
    kl_div = (val1 + val2).mean()
    return z_b, theta_b, dx, x, kl_div


# ----------------------------------------------------------------------------
# a-5 / a-6  RandomFourierEmbedding2d + SpatialGenerator (models.py:53-58,95-123)
# ----------------------------------------------------------------------------
@dataclass
class GeneratorParams:
    coord_w: torch.Tensor                 # (H, E) or (H, 2)
    coord_b: torch.Tensor                 # (H,)
    latent_w: Optional[torch.Tensor]      # (H, zdim)
    hidden_w: Sequence[torch.Tensor]      # L-1 x (H,H)
    hidden_b: Sequence[torch.Tensor]
    out_w: torch.Tensor                   # (n_out, H)
    out_b: torch.Tensor
    fourier_w: Optional[torch.Tensor] = None   # (E,2) buffer
    fourier_b: Optional[torch.Tensor] = None   # (E,)
    sigma: float = 0.01
    activation: str = "leakyrelu"
    resid: bool = False

    @staticmethod
    def from_state_dict(sd, sigma, activation="leakyrelu", resid=False, dtype=torch.float32):
        g = lambda k: sd[k].detach().clone().to(dtype)
        lin = sorted({int(k.split(".")[1]) for k in sd if k.startswith("layers.")})
        pre = (lambda i: f"layers.{i}.linear") if resid else (lambda i: f"layers.{i}")
        hw, hb = [], []
        for i in lin[:-1]:
            hw.append(g(pre(i) + ".weight")); hb.append(g(pre(i) + ".bias"))
        last = lin[-1]
        return GeneratorParams(
            g("coord_linear.weight"), g("coord_linear.bias"),
            g("latent_linear.weight") if "latent_linear.weight" in sd else None,
            hw, hb, g(f"layers.{last}.weight"), g(f"layers.{last}.bias"),
            g("embed_latent.weight") if "embed_latent.weight" in sd else None,
            g("embed_latent.bias") if "embed_latent.bias" in sd else None,
            sigma, activation, resid)

    def named_trainable(self):
        out = [("coord_linear.weight", self.coord_w), ("coord_linear.bias", self.coord_b)]
        if self.latent_w is not None:
            out.append(("latent_linear.weight", self.latent_w))
        for i, (w, b) in enumerate(zip(self.hidden_w, self.hidden_b)):
            out += [(f"hidden{i}.weight", w), (f"hidden{i}.bias", b)]
        out += [("out.weight", self.out_w), ("out.bias", self.out_b)]
        return out


def _act(h, kind):
    return torch.tanh(h) if kind == "tanh" else F.leaky_relu(h, LRELU_SLOPE)


def generator_forward(x, z, p: GeneratorParams):
    """x (B,N,2), z (B,zdim) -> (B,N,n_out)."""
    if x.dim() < 3:
        x = x.unsqueeze(0)
    B, N, _ = x.shape
    f = x.reshape(B * N, 2)
    if p.fourier_w is not None:
        sig = torch.tensor(p.sigma, dtype=torch.float32).to(x.dtype)  # models.py:40
        f = torch.cos(F.linear(f, p.fourier_w / sig, p.fourier_b))    # models.py:57
    h = F.linear(f, p.coord_w, p.coord_b).view(B, N, -1)
    if p.latent_w is not None:
        if z.dim() < 2:
            z = z.unsqueeze(0)
        h = h + F.linear(z, p.latent_w).unsqueeze(1)
    h = _act(h.view(B * N, -1), p.activation)
    for w, b in zip(p.hidden_w, p.hidden_b):
        if p.resid:
            h = _act(F.linear(h, w, b) + h, p.activation)   # ResidLinear, models.py:29-30
        else:
            h = _act(F.linear(h, w, b), p.activation)
    return F.linear(h, p.out_w, p.out_b).view(B, N, -1)


# ----------------------------------------------------------------------------
# a-7 / a-8  likelihoods
# ----------------------------------------------------------------------------
def bernoulli_loglik(y_hat, y):
    """train_mnist.py:288-291: -BCEWithLogits(mean) * size == (1/B) sum_b sum_px."""
    B = y.shape[0]
    y_hat = y_hat.reshape(B, -1)
    y = y.reshape(B, -1)
    size = y.shape[1]
    return -F.binary_cross_entropy_with_logits(y_hat, y) * size


def bernoulli_loglik_rgb(y_hat, y):
    """train_galaxy.py:288-292: compares y_hat.view(B,-1,3) with y.view(B,-1,3);
    y keeps its raw memory order."""
    B = y.shape[0]
    y_hat = y_hat.reshape(B, -1, 3)
    y = y.reshape(B, -1, 3)
    size = y.shape[1] * 3
    return -F.binary_cross_entropy_with_logits(y_hat, y) * size


def particle_mask(dx, s, n: int, radius: int):
    """train_particles.py:309-324: bool (B, n*n)."""
    xi = torch.arange(-n // 2, n // 2, 1, dtype=torch.float64)
    yi = torch.arange(n // 2, -n // 2, -1, dtype=torch.float64)
    yy, xx = torch.meshgrid(yi, xi, indexing="ij")
    gx, gy = xx.reshape(1, -1), yy.reshape(1, -1)
    center = dx.detach().to(torch.float32) / s.to(torch.float32)  # (B,1,2) fp32 division as numpy does
    center = center.to(torch.float64)
    dist = torch.sqrt((center[:, :, 0] - gx) ** 2 + (center[:, :, 1] - gy) ** 2)
    return dist < radius


def gaussian_loglik(y_hat, y, n: int, ctf=None, mask=None):
    """train_particles.py:285-338 (fit_noise with CTF is broken upstream, not replicated)."""
    B = y.shape[0]
    y = y.reshape(B, -1)
    y_hat = y_hat.reshape(B, -1)
    y_mu, y_logvar, y_var = y_hat, None, None
    if y_hat.shape[1] > y.shape[1]:
        # generator output is (B, N, 2) flattened -> interleaved; reference slices flat halves
        y_mu = y_hat[:, :y.shape[1]]
        y_logvar = y_hat[:, y.shape[1]:]
        y_var = torch.exp(y_logvar)
    if ctf is not None:
        pad = ctf.shape[2] // 2
        y_mu = F.conv2d(y_mu.reshape(1, -1, n, n), ctf.to(y_mu.dtype), padding=pad, groups=ctf.shape[0]).reshape(-1, n * n)
    if mask is not None:
        y = torch.where(mask, y, torch.zeros_like(y))
        y_mu = torch.where(mask, y_mu, torch.zeros_like(y_mu))
        if y_var is not None:
            raise NotImplementedError("fit-noise + mask flattens across the batch upstream; out of scope")
    if y_var is not None:
        return -0.5 * torch.sum((y_mu - y) ** 2 / y_var + y_logvar, 1).mean()
    return -0.5 * torch.sum((y_mu - y) ** 2, 1).mean()


# ----------------------------------------------------------------------------
# whole step: eval_minibatch (train_mnist.py:26-294 / train_particles.py:28-343)
# ----------------------------------------------------------------------------
@dataclass
class StepConfig:
    G: int = 8
    padding: int = 8
    rot_refinement: bool = True
    normal_prior_over_r: bool = False
    theta_prior: float = math.pi
    likelihood: str = "bernoulli"      # bernoulli | bernoulli_rgb | gaussian
    mask_radius: int = 0
    encoder: str = "attn_attn"         # attn_attn (--r-inf attention[+offsets]) | attn_unimodal (--r-inf unimodal, --groupconv 0)


def eval_minibatch(x_coord, y, enc: EncoderParams, gen: GeneratorParams, cfg: StepConfig,
                   gumbel, r_z, r_theta, ctf=None):
    """-> (elbo, log_p_x_g_z, kl_div) plus a dict of intermediates."""
    if cfg.encoder == "attn_unimodal":
        # train_mnist.py:88-183: one rotation, N(0, theta_prior) prior on theta (:171)
        enc_out = plainconv_encoder_forward(y, enc, cfg.padding, gumbel)
        z_b, theta_b, dx, x_t, kl_div = attention_posterior(enc_out, x_coord, r_z, r_theta, 1, 0, cfg.theta_prior)
    else:
        enc_out = encoder_forward(y, enc, cfg.G, cfg.padding, cfg.rot_refinement,
                                  cfg.normal_prior_over_r, cfg.theta_prior, gumbel)
        z_b, theta_b, dx, x_t, kl_div = attention_posterior(enc_out, x_coord, r_z, r_theta, cfg.G, cfg.G, cfg.theta_prior)
    y_hat = generator_forward(x_t.contiguous(), z_b, gen)
    n = y.shape[-1]
    if cfg.likelihood == "bernoulli":
        log_p = bernoulli_loglik(y_hat, y)
    elif cfg.likelihood == "bernoulli_rgb":
        log_p = bernoulli_loglik_rgb(y_hat, y)
    else:
        mask = None
        if cfg.mask_radius > 0:
            s = (x_coord[1, 0] - x_coord[0, 0])
            mask = particle_mask(dx, s, n, cfg.mask_radius)
        log_p = gaussian_loglik(y_hat, y, n, ctf, mask)
    elbo = log_p - kl_div
    inter = dict(attn=enc_out[0], q_t_r=enc_out[1], a_sampled=enc_out[3], theta=enc_out[5], z=enc_out[6],
                 z_b=z_b, theta_b=theta_b, dx=dx, x_t=x_t, y_hat=y_hat)
    return elbo, log_p, kl_div, inter


# ----------------------------------------------------------------------------
# a-10  get_latent, attention branch (clustering_mnist.py:122-161)
# ----------------------------------------------------------------------------
def get_latent(x_coord, y, enc: EncoderParams, G: int, padding: int, rot_refinement: bool,
               normal_prior_over_r: bool, theta_prior: float, encoder: str = "attn_attn"):
    dt = y.dtype
    if encoder == "attn_unimodal":     # clustering_mnist.py:81-120 (t_inf attention, r_inf unimodal)
        attn, theta, z = plainconv_head_maps(y, enc, padding)
    else:
        attn_raw, theta, z = encoder_head_maps(y, enc, G, padding)
        attn = attn_raw + rotation_log_prior(G, rot_refinement, normal_prior_over_r, theta_prior, dt)
        if rot_refinement:
            theta = torch.stack((theta[:, 0] + rotation_offsets(G, True, dt).view(1, G, 1, 1), theta[:, 1]), 1)
    B, R, d, _ = attn.shape
    ind = attn.reshape(B, -1).argmax(1)
    ar = torch.arange(B)
    zv = z.reshape(B, z.shape[1], -1)
    zd = zv.shape[1] // 2
    z_content = torch.cat((zv[ar, :zd, ind], torch.exp(zv[ar, zd:, ind])), dim=1)
    theta_mu = theta.reshape(B, 2, -1)[ar, 0:1, ind]
    sm = F.softmax(attn.reshape(B, -1), dim=1).view(B, R, d * d).sum(1)
    s = (x_coord[1, 0] - x_coord[0, 0]).to(dt)
    dx = sm @ translation_grid(d, s, dt)
    return z_content, theta_mu, dx, ind


# ----------------------------------------------------------------------------
# helpers shared by tests / bench
# ----------------------------------------------------------------------------
def image_coords(n: int, dtype=torch.float32):
    """x_coord of train_mnist.py:474-479 (float64 linspace -> float32)."""
    import numpy as np
    xg = np.linspace(-1, 1, n)
    yg = np.linspace(1, -1, n)
    x0, x1 = np.meshgrid(xg, yg)
    return torch.from_numpy(np.stack([x0.ravel(), x1.ravel()], 1)).float().to(dtype)
