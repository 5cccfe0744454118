"""GPU parity of the tcgen05 GEMM core (NT and TN forms, the product's LinearNT / LinearTN kernels) against an fp64
torch reference.

Operands are FP16 (tcgen05.mma.kind::f16: 11-bit significand, the precision class of TF32), products are exact in fp32
and accumulated in fp32.  Two checks: against fp64 on the fp32 source values (only the fp16 rounding of the operands,
2^-11 relative per element: < 1e-3 relative Frobenius) and against fp64 on the rounded operands (fp32 accumulation
only: < 2e-5).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


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


# K (row pitch) must be a multiple of 8 halves: TMA rows are 16-byte aligned
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 64), (256, 128, 128), (300, 512, 104),
                                   (1521, 1024, 784), (5000, 16, 128), (77, 260, 40)])
def test_linear_nt(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    Ah, Bh = A.half(), B.half()
    C = _nt(Ah, Bh)
    ref = A.double() @ B.double().t()
    ref_h = Ah.double() @ Bh.double().t()
    e_full = float((C.double() - ref).norm() / ref.norm())
    e_round = float((C.double() - ref_h).norm() / ref_h.norm())
    print(f"NT {M}x{N}x{K}: rel err vs fp64 {e_full:.3e}, vs fp16-rounded operands {e_round:.3e}")
    assert e_full < 1e-3
    assert e_round < 2e-5


def test_linear_nt_bias_act():
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(700, 96, device="cuda", generator=g).half()
    B = torch.randn(384, 96, device="cuda", generator=g).half()
    bias = torch.randn(384, device="cuda", generator=g)
    C = _nt(A, B, bias, act=1)
    ref = torch.nn.functional.leaky_relu(A.double() @ B.double().t() + bias.double(), 0.01)
    assert float((C.double() - ref).norm() / ref.norm()) < 2e-5


@pytest.mark.parametrize("R,Ma,Nb", [(64, 128, 128), (64, 128, 256), (1000, 128, 128), (5000, 512, 512),
                                     (20000, 1024, 800), (333, 104, 40)])
@pytest.mark.parametrize("transpose_out", [0, 1])
def test_linear_tn(R, Ma, Nb, transpose_out):
    g = torch.Generator(device="cuda").manual_seed(R + Ma + Nb)
    P = torch.randn(R, Ma, device="cuda", generator=g).half()
    Q = torch.randn(R, Nb, device="cuda", generator=g).half()
    C = _tn(P, Q, transpose_out)
    ref = P.double().t() @ Q.double()
    if transpose_out:
        ref = ref.t()
    e = float((C.double() - ref).norm() / ref.norm())
    print(f"TN R={R} {Ma}x{Nb} T={transpose_out}: rel err vs fp16-rounded operands {e:.3e}")
    assert e < 5e-5


def _nt_full(A, B, bias=None, act=0, aux=None, acc_scale=None, store_scale=None, colsum=False, proj=None):
    """-> (C16 as float, colsum or None, proj_out or None) of the hidden-layer GEMM with its whole epilogue"""
    import ctypes
    from tvae_b200 import _lib
    M, K = A.shape
    N = B.shape[0]
    C16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    cs = torch.zeros(N, device="cuda") if colsum else None
    pw, pb, po, n_proj = None, None, None, 0
    if proj is not None:
        pw, pb = proj
        n_proj = pw.shape[0]
        po = torch.zeros(M, n_proj, device="cuda")
    fn = _lib.lib().tvae_test_linear_nt_full
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    rc = fn(_lib.ptr(A), _lib.ptr(B), M, N, K, _lib.ptr(bias), act, _lib.ptr(C16), _lib.ptr(aux), 0, _lib.ptr(acc_scale),
            _lib.ptr(store_scale), _lib.ptr(cs), _lib.ptr(pw), _lib.ptr(pb), _lib.ptr(po), n_proj, _lib.stream_ptr())
    _lib.check(rc, "tvae_test_linear_nt_full")
    torch.cuda.synchronize()
    return C16.float(), cs, po


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 256), (1000, 512, 512), (70001, 512, 512), (4097, 512, 128), (5, 256, 512)])
@pytest.mark.parametrize("mode", ["forward", "dgrad"])
def test_linear_nt_full_epilogue(M, N, K, mode):
    """LinearNT with the whole epilogue of the generator's hidden layers - forward: bias + LeakyReLU + fp16 TMA stores + fused
    output projection; input gradient: derivative mask, power-of-two scales, column sums - against fp64 on the fp16-rounded
    operands."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    if mode == "forward":
        bias = torch.randn(N, device="cuda", generator=g) * 0.1
        pw = torch.randn(3, N, device="cuda", generator=g) * 0.1
        pb = torch.randn(3, device="cuda", generator=g)
        C, _, po = _nt_full(A, B, bias=bias, act=1, proj=(pw, pb))
        ref = torch.nn.functional.leaky_relu(A.double() @ B.double().t() + bias.double(), 0.01)
        ref_proj = ref @ pw.double().t() + pb.double()
        assert float((C.double() - ref).norm() / ref.norm()) < 6e-4          # fp16 rounding of the stored output
        assert float((po.double() - ref_proj).norm() / ref_proj.norm()) < 2e-5
    else:
        aux = torch.randn(M, N, device="cuda", generator=g).half()
        acc_scale = torch.tensor([0.25], device="cuda")
        store_scale = torch.tensor([8.0], device="cuda")
        C, cs, _ = _nt_full(A, B, aux=aux, acc_scale=acc_scale, store_scale=store_scale, colsum=True)
        mask = torch.where(aux.double() > 0, 1.0, 0.01)
        val = 0.25 * (A.double() @ B.double().t()) * mask
        assert float((C.double() / 8.0 - val).norm() / val.norm()) < 6e-4
        assert float((cs.double() - val.sum(0)).norm() / val.sum(0).norm()) < 1e-4


@pytest.mark.parametrize("R,Ma,Nb", [(5000, 512, 512), (70001, 512, 512), (4096, 256, 512), (999, 512, 256), (3000, 200, 160)])
def test_linear_tn_pair_matches_tc_gemm(R, Ma, Nb):
    """The CTA-pair weight-gradient kernel against the tc_gemm LinearTN policy it replaces for wide outputs."""
    from tvae_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(R + Ma + Nb)
    P = torch.randn(R, Ma, device="cuda", generator=g).half()
    Q = torch.randn(R, Nb, device="cuda", generator=g).half()
    ref = P.double().t() @ Q.double()
    outs = []
    try:
        for pair in (1, 0):
            _lib.lib().tvae_test_set_fast_paths(pair)
            outs.append([_tn(P, Q, t) for t in (0, 1)])
    finally:
        _lib.lib().tvae_test_set_fast_paths(1)
    for o in outs:
        assert float((o[0].double() - ref).norm() / ref.norm()) < 5e-5
        assert float((o[1].double() - ref.t()).norm() / ref.norm()) < 5e-5
