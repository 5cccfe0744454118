"""GPU parity of the tcgen05 GEMM core (NT and TN forms) against an fp64 torch reference.

Tolerance: operands are consumed as TF32 (10-bit mantissa, truncated by the tensor core), products are
accumulated in fp32: |err| <= ~2^-10 * sum|a||b| worst case; we assert relative Frobenius error < 2e-3
and additionally compare against a reference fed with TF32-truncated operands (< 2e-5).
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _trunc_tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _nt(A, B, bias=None, act=0):
    from tvae_b200 import _lib
    M, K = A.shape
    N = B.shape[0]
    C = torch.empty(M, N, device="cuda", dtype=torch.float32)
    rc = _lib.lib().tvae_test_linear_nt(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, _lib.ptr(bias), act,
                                        _lib.stream_ptr())
    _lib.check(rc, "tvae_test_linear_nt")
    torch.cuda.synchronize()
    return C


def _tn(P, Q, transpose_out=0):
    from tvae_b200 import _lib
    R, Ma = P.shape
    Nb = Q.shape[1]
    C = torch.zeros((Nb, Ma) if transpose_out else (Ma, Nb), device="cuda", dtype=torch.float32)
    rc = _lib.lib().tvae_test_linear_tn(_lib.ptr(P), _lib.ptr(Q), _lib.ptr(C), R, Ma, Nb, transpose_out,
                                        _lib.stream_ptr())
    _lib.check(rc, "tvae_test_linear_tn")
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 256, 64), (256, 128, 128), (300, 512, 100),
                                   (1521, 1024, 784), (5000, 16, 128), (77, 260, 36)])
def test_linear_nt(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    C = _nt(A, B)
    ref = A.double() @ B.double().t()
    ref_t = _trunc_tf32(A).double() @ _trunc_tf32(B).double().t()
    e_full = float((C.double() - ref).norm() / ref.norm())
    e_trunc = float((C.double() - ref_t).norm() / ref_t.norm())
    print(f"NT {M}x{N}x{K}: rel err vs fp64 {e_full:.3e}, vs tf32-truncated operands {e_trunc:.3e}")
    assert e_full < 2e-3
    assert e_trunc < 2e-5


def test_linear_nt_bias_act():
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(700, 96, device="cuda", generator=g)
    B = torch.randn(384, 96, device="cuda", generator=g)
    bias = torch.randn(384, device="cuda", generator=g)
    C = _nt(A, B, bias, act=1)
    ref = torch.nn.functional.leaky_relu(_trunc_tf32(A).double() @ _trunc_tf32(B).double().t() + bias.double(), 0.01)
    assert float((C.double() - ref).norm() / ref.norm()) < 2e-5


@pytest.mark.parametrize("R,Ma,Nb", [(32, 128, 128), (64, 128, 256), (1000, 128, 128), (5000, 512, 512),
                                     (20000, 1024, 800), (333, 100, 40)])
@pytest.mark.parametrize("transpose_out", [0, 1])
def test_linear_tn(R, Ma, Nb, transpose_out):
    g = torch.Generator(device="cuda").manual_seed(R + Ma + Nb)
    P = torch.randn(R, Ma, device="cuda", generator=g)
    Q = torch.randn(R, Nb, device="cuda", generator=g)
    C = _tn(P, Q, transpose_out)
    ref_t = _trunc_tf32(P).double().t() @ _trunc_tf32(Q).double()
    if transpose_out:
        ref_t = ref_t.t()
    e = float((C.double() - ref_t).norm() / ref_t.norm())
    print(f"TN R={R} {Ma}x{Nb} T={transpose_out}: rel err vs tf32-truncated operands {e:.3e}")
    assert e < 5e-5
