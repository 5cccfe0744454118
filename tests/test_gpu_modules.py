"""Standalone forwards of the drop-in module classes (calls a user of the reference can make outside the fused step):
RandomFourierEmbedding2d.forward (models.py:53-58), ResidLinear.forward (models.py:29-30) and the module-interface tail
q_t_r / a_sampled of the encoder (models.py:383-388), forward and backward, against torch in fp64.
Tolerances: the Fourier embedding and the softmax pair are fp32 kernels (1e-5); ResidLinear runs on the tensor cores with
FP16 operands (2^-11 relative per operand element: 2e-3 forward, 5e-3 gradients)."""
import contextlib
import io

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_random_fourier_embedding_forward_backward():
    import src.models as models
    with contextlib.redirect_stdout(io.StringIO()):
        emb = models.RandomFourierEmbedding2d(2, 1024, sigma=2.0 / 49).to(DEV)
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(2, 300, 2, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    out = emb(x)
    assert tuple(out.shape) == (2, 300, 1024)
    xd = x.detach().double().cpu().requires_grad_(True)
    ref = torch.cos(F.linear(xd, emb.weight.double().cpu() / torch.tensor(2.0 / 49, dtype=torch.float32).double(), emb.bias.double().cpu()))
    # the phase reaches ~50 rad: fp32 evaluation of cos(phase) carries ~phase * 6e-8 absolute error
    assert float((out.detach().cpu().double() - ref).abs().max()) < 2e-5
    w = torch.randn(out.shape, generator=g)
    (out * w.to(DEV)).sum().backward()
    (ref * w.double()).sum().backward()
    assert rel_err(x.grad.cpu(), xd.grad) < 1e-4
    assert emb(None) == 0                                                    # models.py:54-55


@pytest.mark.parametrize("act", [nn.LeakyReLU, nn.Tanh])
@pytest.mark.parametrize("M,H", [(700, 64), (4096, 512)])
def test_resid_linear_forward_backward(act, M, H):
    import src.models as models
    torch.manual_seed(1)
    layer = models.ResidLinear(H, H, activation=act).to(DEV)
    x = torch.randn(M, H, device=DEV, requires_grad=True)
    y = layer(x)
    xd = x.detach().double().cpu().requires_grad_(True)
    wd, bd = layer.linear.weight.detach().double().cpu().requires_grad_(True), layer.linear.bias.detach().double().cpu().requires_grad_(True)
    a = (lambda t: F.leaky_relu(t, 0.01)) if act is nn.LeakyReLU else torch.tanh
    ref = a(F.linear(xd, wd, bd) + xd)
    assert rel_err(y.detach().cpu(), ref.detach()) < 2e-3
    w = torch.randn(M, H)
    (y * w.to(DEV)).sum().backward()
    # same activation pattern for the derivative (a LeakyReLU unit within fp16 rounding of zero may flip)
    pat = y.detach().double().cpu()
    dact = torch.where(pat > 0, 1.0, 0.01) if act is nn.LeakyReLU else 1 - pat ** 2
    dpre = w.double() * dact
    ref_dx = dpre @ (wd.detach() + torch.eye(H, dtype=torch.float64))
    ref_dw = dpre.t() @ xd.detach()
    assert rel_err(x.grad.cpu(), ref_dx) < 5e-3
    assert rel_err(layer.linear.weight.grad.cpu(), ref_dw) < 5e-3
    assert rel_err(layer.linear.bias.grad.cpu(), dpre.sum(0)) < 1e-4


@pytest.mark.parametrize("C,n,k,p,G,O", [(1, 14, 7, 3, 4, 32), (3, 10, 6, 2, 8, 32)])
def test_groupconv_input_gradient(C, n, k, p, G, O):
    """GroupConv used on its own with an input that requires a gradient (models.py:202-225 under autograd): the
    CUDA-core input-gradient kernel against autograd through the fp64 oracle's group convolution.  The forward runs
    FP16 operands (2e-3); the input gradient is an fp32 kernel over the fp32 rotated bank (1e-4)."""
    import src.models as models
    from oracle import target_vae_oracle as orc
    torch.manual_seed(2)
    conv = models.GroupConv(C, O, k, stride=1, padding=p, bias=True, input_rot_dim=1, output_rot_dim=G).to(DEV)
    y = torch.randn(2, C, n, n, device=DEV, requires_grad=True)
    out = conv(y, DEV)
    yd = y.detach().double().cpu().requires_grad_(True)
    wd = conv.weight.detach().double().cpu().requires_grad_(True)
    ref = orc.groupconv_forward(yd, wd, conv.bias.detach().double().cpu(), G, p)
    assert tuple(out.shape) == tuple(ref.shape)
    assert rel_err(out.detach().cpu(), ref.detach()) < 2e-3
    w = torch.randn(ref.shape)
    (out * w.to(DEV)).sum().backward()
    (ref * w.double()).sum().backward()
    assert rel_err(y.grad.cpu(), yd.grad) < 1e-4
    assert rel_err(conv.weight.grad.cpu(), wd.grad) < 5e-3


def test_softmax_pair_autograd():
    from tvae_b200 import functional as TF
    g = torch.Generator().manual_seed(5)
    B, shape = 3, (3, 4, 9, 9)
    attn = torch.randn(shape, generator=g).to(DEV).requires_grad_(True)
    gum = torch.randn(B, 4 * 81, generator=g).to(DEV)
    q, a = TF.SoftmaxPairFn.apply(attn, gum)
    ad = attn.detach().double().cpu().requires_grad_(True)
    qr = F.log_softmax(ad.reshape(B, -1), 1).view(shape)
    ar = F.softmax(ad.reshape(B, -1) + gum.double().cpu(), 1).view(shape)
    assert rel_err(q.detach().cpu(), qr.detach()) < 1e-5 and rel_err(a.detach().cpu(), ar.detach()) < 1e-5
    wq, wa = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    ((q * wq.to(DEV)).sum() + (a * wa.to(DEV)).sum()).backward()
    ((qr * wq.double()).sum() + (ar * wa.double()).sum()).backward()
    assert rel_err(attn.grad.cpu(), ad.grad) < 1e-5
    # only one of the two outputs used
    attn.grad = None
    q2, _ = TF.SoftmaxPairFn.apply(attn, gum)
    (q2 * wq.to(DEV)).sum().backward()
    ad.grad = None
    (F.log_softmax(ad.reshape(B, -1), 1).view(shape) * wq.double()).sum().backward()
    assert rel_err(attn.grad.cpu(), ad.grad) < 1e-5


def test_encoder_module_tail_carries_gradients():
    """The 7-tuple's q_t_r / a_sampled come out of SoftmaxPairFn: a loss on them reaches the encoder parameters."""
    from test_gpu_step import build_models
    from tvae_b200.config import HotPathConfig
    cfg = HotPathConfig("tail", C=1, n=20, k=9, p=3, G=4, z=2, O=32, hidden=32)
    _, enc = build_models(cfg)
    y = torch.rand(2, 1, cfg.n, cfg.n, device=DEV)
    attn, q, p_r, a_s, offs, theta, z = enc(y, DEV)
    assert abs(float(a_s.sum()) - 2.0) < 1e-4 and abs(float(q.exp().sum()) - 2.0) < 1e-4
    ((a_s * attn.detach()).sum() + q.mean()).backward()
    assert float(enc.conv1.weight.grad.abs().max()) > 0
