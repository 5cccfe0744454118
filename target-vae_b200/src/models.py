"""Drop-in `src.models` for TARGET-VAE's hot path, backed by the sm_100a kernels in libtvae_b200.so.

Same class names, constructor signatures, parameter names/shapes (state_dict compatible with the reference,
SURVEY.md §8b) and `forward` signatures/returns as the reference's src/models.py, so that
`train_*.py` / `clustering_*.py` can import this file unchanged and reference checkpoints (whole-module
pickles resolved as `src.models.<Name>`) load.  The arithmetic is not PyTorch: forward/backward of
GroupConv, the attention encoder and the spatial generator run hand-written CUDA (tcgen05 FP16-operand GEMMs with
operand generators, fused heads) through the C ABI in include/tvae_b200.h.  There is no CPU fallback: calling
`forward` on CPU tensors or without the built library raises.

Covered variants: LeakyReLU and tanh activations, residual generator layers, the attention /
unimodal-rotation encoder with a plain convolution (groupconv = 0) or a rotation-pooled group convolution
(groupconv > 0).  The MLP encoder of the unimodal/unimodal spatial-VAE baseline (not in SURVEY.md §8) is importable,
holds the right parameters and state_dict, and its `forward` raises.
"""
from __future__ import print_function, division

import math

import numpy as np
import torch
import torch.nn as nn
from torch.nn import Parameter
from torch.nn.modules.utils import _pair

from tvae_b200 import functional as TF
from tvae_b200 import ops as _ops


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(what + ": the TARGET-VAE hot path runs on sm_100a only (no CPU fallback); move the "
                           "module and its inputs to a CUDA device")


class ResidLinear(nn.Module):
    """models.py:22-30: act(linear(x) + x).  Inside SpatialGenerator it runs on the hidden-layer tensor-core kernels as a
    plain layer with the effective weight W + I (tvae_b200.functional._gen_weights); called on its own it runs the same
    way through tvae_linear_act_fwd / _bwd (FP16 operands, fp32 accumulation)."""

    def __init__(self, n_in, n_out, activation=nn.LeakyReLU):
        super(ResidLinear, self).__init__()
        self.linear = nn.Linear(n_in, n_out)
        self.act = activation()

    def forward(self, x):
        _require_cuda(x, "ResidLinear.forward")
        return TF.LinearActFn.apply(x, self.linear.weight, self.linear.bias, True, _ops.act_kind(self.act))


class RandomFourierEmbedding2d(nn.Module):
    """models.py:33-58: buffers `weight` ~ randn(E,2), `bias` ~ U(0, 2pi); cos(x W^T / sigma + b).
    Inside SpatialGenerator the expansion is generated tile-by-tile in shared memory by the layer-1 GEMM and never
    materialised; called on its own it is one elementwise kernel (tvae_fourier_embed_fwd / _bwd)."""

    def __init__(self, in_dim, embedding_dim, sigma=0.01):
        super(RandomFourierEmbedding2d, self).__init__()
        self.in_dim = in_dim
        self.embedding_dim = embedding_dim
        self.sigma = torch.tensor(sigma, dtype=torch.float32)
        self.register_buffer('weight', torch.randn(embedding_dim, in_dim))
        self.register_buffer('bias', torch.rand(embedding_dim) * 2 * np.pi)
        print('# sigma value is {}'.format(self.sigma))

    def forward(self, x):
        if x is None:
            return 0
        _require_cuda(x, "RandomFourierEmbedding2d.forward")
        return TF.FourierEmbedFn.apply(x, self.weight, self.bias, float(self.sigma))


class SpatialGenerator(nn.Module):
    """models.py:65-123."""

    def __init__(self, latent_dim, hidden_dim, n_out=1, num_layers=1, activation=nn.LeakyReLU,
                 resid=False, fourier_expansion=False, sigma=0.01):
        super(SpatialGenerator, self).__init__()
        self.fourier_expansion = fourier_expansion
        in_dim = 2
        if fourier_expansion:
            embedding_dim = 1024
            self.embed_latent = RandomFourierEmbedding2d(in_dim, embedding_dim, sigma)
            in_dim = embedding_dim
        self.coord_linear = nn.Linear(in_dim, hidden_dim)
        self.latent_dim = latent_dim
        if latent_dim > 0:
            self.latent_linear = nn.Linear(latent_dim, hidden_dim, bias=False)
        layers = [activation()]
        for _ in range(1, num_layers):
            if resid:
                layers.append(ResidLinear(hidden_dim, hidden_dim, activation=activation))
            else:
                layers.append(nn.Linear(hidden_dim, hidden_dim))
                layers.append(activation())
        layers.append(nn.Linear(hidden_dim, n_out))
        self.layers = nn.Sequential(*layers)

    # ---- helpers used by the fused step --------------------------------------------------------
    # Derived from state the reference's own modules carry, never from attributes only this __init__ would create: the
    # reference checkpoints are whole-module pickles (src/utils.py:42-46) and unpickling does not run __init__.
    @property
    def _sigma(self) -> float:
        """Fourier scale: RandomFourierEmbedding2d.sigma (models.py:40), a 0-d fp32 tensor attribute of the sub-module."""
        return float(self.embed_latent.sigma) if self.fourier_expansion else 0.01

    @property
    def _resid(self) -> bool:
        """--generator-resid-layers: the hidden layers are ResidLinear modules (models.py:84-86)."""
        return any(isinstance(m, ResidLinear) for m in self.layers)

    def _check_supported(self):
        if not hasattr(self, 'latent_linear'):
            raise NotImplementedError("SpatialGenerator: only latent-conditioned generators (latent_dim > 0) are on the "
                                      "accelerated path")
        return _ops.act_kind(self.layers[0])

    def act_kind(self):
        """ops.ACT_* of --activation (LeakyReLU or tanh, train_mnist.py:516-519)."""
        return self._check_supported()

    def hot_path_params(self):
        """coord_linear.{weight,bias}, latent_linear.weight, (hidden weight, bias)*, out weight, bias."""
        self._check_supported()
        lin = [m.linear if isinstance(m, ResidLinear) else m for m in self.layers if isinstance(m, (nn.Linear, ResidLinear))]
        ps = [self.coord_linear.weight, self.coord_linear.bias, self.latent_linear.weight]
        for m in lin:
            ps += [m.weight, m.bias]
        return ps

    def fourier_buffers(self):
        if self.fourier_expansion:
            return self.embed_latent.weight, self.embed_latent.bias
        return None, None

    def forward(self, x, z):
        if len(x.size()) < 3:
            x = x.unsqueeze(0)
        if len(z.size()) < 2:
            z = z.unsqueeze(0)
        _require_cuda(x, "SpatialGenerator.forward")
        fw, fb = self.fourier_buffers()
        return TF.GeneratorFn.apply(fw, fb, (self._sigma, self._resid, self.act_kind()), x, z, *self.hot_path_params())


class GroupConv(nn.Module):
    """models.py:132-225: P_G lifting convolution.  The rotated filter bank is sampled once per call into an fp16 operand
    that stays L2-resident (8-25 MB), the im2col operand is generated in shared memory, the contraction runs on the tcgen05
    tensor cores."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1,
                 padding=0, bias=True, input_rot_dim=1, output_rot_dim=4):
        super(GroupConv, self).__init__()
        self.ksize = kernel_size
        kernel_size = _pair(kernel_size)
        stride = _pair(stride)
        padding = _pair(padding)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.input_rot_dim = input_rot_dim
        self.output_rot_dim = output_rot_dim
        self.weight = Parameter(torch.Tensor(out_channels, in_channels, self.input_rot_dim, *kernel_size), requires_grad=True)
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels), requires_grad=True)
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def trans_filter(self, device):
        """(O, G, C, 1, k, k) rotated copies of the base filter (models.py:174-197)."""
        _require_cuda(self.weight, "GroupConv.trans_filter")
        O, C, _, k, _ = self.weight.shape
        s = _ops.enc_shape(1, C, k, k, 0, self.output_rot_dim, O, 1)
        bank = _ops.filter_bank_fwd(s, self.weight)[:, :C * k * k].float()   # stored fp16 (GEMM operand)
        return bank.view(self.output_rot_dim, O, C, 1, k, k).permute(1, 0, 2, 3, 4, 5)

    def forward(self, input, device):
        _require_cuda(input, "GroupConv.forward")
        if self.stride != (1, 1) or self.input_rot_dim != 1 or self.padding[0] != self.padding[1]:
            raise NotImplementedError("GroupConv: only stride 1, square padding, input_rot_dim 1 are on the hot path")
        return TF.GroupConvFn.apply(input, self.weight, self.bias, self.output_rot_dim, self.padding[0])


class InferenceNetwork_UnimodalTranslation_UnimodalRotation(nn.Module):
    """models.py:229-260 (spatial-VAE style MLP baseline, outside SURVEY.md §8: parameters / state_dict only)."""

    def __init__(self, n, latent_dim, hidden_dim, num_layers=1, activation=nn.LeakyReLU, resid=False):
        super(InferenceNetwork_UnimodalTranslation_UnimodalRotation, self).__init__()
        self.latent_dim = latent_dim
        self.n = n
        print('n is {}'.format(n))
        layers = [nn.Linear(n, hidden_dim), activation()]
        for _ in range(1, num_layers):
            if resid:
                layers.append(ResidLinear(hidden_dim, hidden_dim, activation=activation))
            else:
                layers.append(nn.Linear(hidden_dim, hidden_dim))
                layers.append(activation())
        layers.append(nn.Linear(hidden_dim, 2 * latent_dim))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        raise NotImplementedError("unimodal/unimodal inference (the spatial-VAE MLP baseline) is outside the accelerated hot path")


class InferenceNetwork_AttentionTranslation_UnimodalRotation(nn.Module):
    """models.py:268-319: attention over translations only.  groupconv = 0 (plain Conv2d(C, O, n, padding n//2) ->
    LeakyReLU -> 1x1 conv -> LeakyReLU -> heads) runs on the same tcgen05 kernels as the TARGET-VAE encoder with a
    single, unrotated filter slot (G = 1); groupconv > 0 is the P_G group conv followed by the rotation pooling `fc_r`
    (rot_pool_fwd / rot_pool_bwd kernels) and then one rotation slot.  forward returns the reference's 4-tuple."""

    def __init__(self, n, in_channels, latent_dim, kernels_num=128, activation=nn.LeakyReLU, groupconv=0):
        super(InferenceNetwork_AttentionTranslation_UnimodalRotation, self).__init__()
        self.activation = activation()
        self.latent_dim = latent_dim
        self.input_size = n
        self.kernels_num = kernels_num
        self.groupconv = groupconv
        if self.groupconv == 0:
            self.conv1 = nn.Conv2d(in_channels, self.kernels_num, self.input_size, padding=self.input_size // 2)
            self.conv2 = nn.Conv2d(self.kernels_num, self.kernels_num, 1)
        else:
            self.conv1 = GroupConv(in_channels, self.kernels_num, self.input_size, padding=self.input_size // 2,
                                   input_rot_dim=1, output_rot_dim=self.groupconv)
            self.conv2 = nn.Conv2d(self.kernels_num, self.kernels_num, 1)
            self.fc_r = nn.Linear(self.groupconv, 1)
        self.conv_a = nn.Conv2d(self.kernels_num, 1, 1)
        self.conv_r = nn.Conv2d(self.kernels_num, 2, 1)
        self.conv_z = nn.Conv2d(self.kernels_num, 2 * self.latent_dim, 1)

    # geometry in the TARGET-VAE encoder's terms (used by tvae_b200.elbo): properties of state the reference module
    # also has, so that a reference checkpoint (a whole-module pickle, restored without __init__) works unchanged
    @property
    def kernels_size(self) -> int:
        return self.input_size

    @property
    def padding(self) -> int:
        return self.input_size // 2

    # ---- helpers used by the fused step --------------------------------------------------------
    def encoder_spec(self, theta_prior=np.pi):
        """One rotation slot, no rotation prior, no offsets; the N(0, theta_prior) prior on theta is the trainer's
        argument (train_mnist.py:171), not a module attribute."""
        pool = self.groupconv > 0        # group conv over G rotations, pooled to one slot by fc_r (models.py:301-304)
        return TF.EncoderSpec(self.groupconv if pool else 1, self.padding, self.latent_dim, False, False, float(theta_prior),
                              theta_prior_std=float(theta_prior), act=_ops.act_kind(self.activation), pool=pool)

    def hot_path_params(self):
        ps = [self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, self.conv_a.weight, self.conv_a.bias,
              self.conv_r.weight, self.conv_r.bias, self.conv_z.weight, self.conv_z.bias]
        if self.groupconv > 0:
            ps += [self.fc_r.weight, self.fc_r.bias]
        return ps

    def head_maps(self, x):
        """(B, 3+2z, 1, H', W') = [attn, theta_mu, theta_logstd, z...] with autograd."""
        _require_cuda(x, "encoder forward")
        return TF.EncoderHeadsFn.apply(self.encoder_spec(), x, *self.hot_path_params())

    def forward(self, x, device):
        heads = self.head_maps(x).squeeze(2)
        B = heads.shape[0]
        attn = heads[:, 0:1]
        theta = heads[:, 1:3]
        z = heads[:, 3:]
        # models.py:311-313: Gumbel-softmax sample of the attention map (tvae_attn_softmax_pair; the fused step never
        # materialises it)
        _, a_sampled = TF.SoftmaxPairFn.apply(attn.reshape(B, -1), TF.gumbel_noise((B, heads.shape[2] * heads.shape[3]), heads.device))
        return attn, a_sampled.view(B, heads.shape[2], heads.shape[3]), theta, z


class InferenceNetwork_AttentionTranslation_AttentionRotation(nn.Module):
    """models.py:326-403: the TARGET-VAE encoder.  conv1 (P_G group conv) -> LeakyReLU -> 1x1x1 conv ->
    LeakyReLU -> attention / theta / z heads run as two tcgen05 kernels; forward returns the reference's 7-tuple."""

    def __init__(self, n, in_channels, latent_dim, kernels_num=128, kernels_size=65, padding=16, activation=nn.LeakyReLU,
                 groupconv=0, rot_refinement=False, theta_prior=np.pi, normal_prior_over_r=True):
        super(InferenceNetwork_AttentionTranslation_AttentionRotation, self).__init__()
        self.activation = activation()
        self.latent_dim = latent_dim
        self.input_size = n
        self.kernels_num = kernels_num
        self.kernels_size = kernels_size
        self.padding = padding
        self.groupconv = groupconv
        self.rot_refinement = rot_refinement
        self.theta_prior = theta_prior
        self.normal_prior_over_r = normal_prior_over_r
        self.conv1 = GroupConv(in_channels, self.kernels_num, self.kernels_size, padding=self.padding, input_rot_dim=1,
                               output_rot_dim=self.groupconv)
        self.conv2 = nn.Conv3d(self.kernels_num, self.kernels_num, 1)
        self.conv_a = nn.Conv3d(self.kernels_num, 1, 1)
        self.conv_r = nn.Conv3d(self.kernels_num, 2, 1)
        self.conv_z = nn.Conv3d(self.kernels_num, 2 * self.latent_dim, 1)

    # ---- helpers used by the fused step --------------------------------------------------------
    def encoder_spec(self):
        if self.groupconv < 1:
            raise NotImplementedError("the accelerated encoder needs groupconv in {4, 8, 16}")
        return TF.EncoderSpec(self.groupconv, self.padding, self.latent_dim, bool(self.rot_refinement),
                              bool(self.normal_prior_over_r), float(self.theta_prior), act=_ops.act_kind(self.activation))

    def hot_path_params(self):
        return [self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, self.conv_a.weight, self.conv_a.bias,
                self.conv_r.weight, self.conv_r.bias, self.conv_z.weight, self.conv_z.bias]

    def head_maps(self, x):
        """(B, 3+2z, G, H', W') = [attn + p_r, theta_mu + offset, theta_logstd, z...] with autograd."""
        _require_cuda(x, "encoder forward")
        return TF.EncoderHeadsFn.apply(self.encoder_spec(), x, *self.hot_path_params())

    def forward(self, x, device):
        spec = self.encoder_spec()
        heads = self.head_maps(x)
        B = heads.shape[0]
        attn = heads[:, 0]
        theta = heads[:, 1:3]
        z = heads[:, 3:]
        p_r_list, offs_list = spec.tables()
        p_r = torch.tensor(p_r_list, dtype=torch.float32, device=heads.device).unsqueeze(1).unsqueeze(2)
        offsets = torch.tensor(offs_list, dtype=torch.float32, device=heads.device)
        # module-interface tail (models.py:383-388): log_softmax and the Gumbel-softmax sample over all (r, t) cells in one
        # kernel (tvae_attn_softmax_pair, backward tvae_attn_softmax_pair_bwd); the fused training step (tvae_b200.elbo)
        # never materialises these.
        q_t_r, a_sampled = TF.SoftmaxPairFn.apply(attn, TF.gumbel_noise((B, attn[0].numel()), heads.device))
        return attn, q_t_r, p_r, a_sampled, offsets, theta, z
