"""Bring-up probe for the tcgen05 GEMM core: correctness summary + rough throughput (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "target-vae_b200"))
import torch
from tvae_b200 import _lib

L = _lib.lib()
def trunc(x): return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)

def nt(A, B, C):
    _lib.check(L.tvae_test_linear_nt(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), A.shape[0], B.shape[0], A.shape[1], None, 0, _lib.stream_ptr()), "nt")
def tn(P, Q, C, t=0):
    _lib.check(L.tvae_test_linear_tn(_lib.ptr(P), _lib.ptr(Q), _lib.ptr(C), P.shape[0], P.shape[1], Q.shape[1], t, _lib.stream_ptr()), "tn")

torch.manual_seed(0)
for (M, N, K) in [(128, 128, 32), (128, 256, 32), (256, 256, 64), (1521, 1024, 784)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.full((M, N), float("nan"), device="cuda")
    try:
        nt(A, B, C); torch.cuda.synchronize()
        ref = trunc(A).double() @ trunc(B).double().t()
        print(f"NT {M}x{N}x{K} rel err {float((C.double()-ref).norm()/ref.norm()):.3e} nan={int(torch.isnan(C).sum())}", flush=True)
    except Exception as e:
        print("NT failed", (M, N, K), e, flush=True); break
for (R, Ma, Nb) in [(32, 128, 128), (64, 128, 256), (4096, 512, 512)]:
    P = torch.randn(R, Ma, device="cuda"); Q = torch.randn(R, Nb, device="cuda"); C = torch.zeros(Ma, Nb, device="cuda")
    try:
        tn(P, Q, C); torch.cuda.synchronize()
        ref = trunc(P).double().t() @ trunc(Q).double()
        print(f"TN R={R} {Ma}x{Nb} rel err {float((C.double()-ref).norm()/ref.norm()):.3e}", flush=True)
    except Exception as e:
        print("TN failed", (R, Ma, Nb), e, flush=True); break

def bench(fn, flops, name, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: {ms:.3f} ms  {flops/ms/1e9:.1f} TFLOP/s", flush=True)

M, N, K = 152064, 1024, 800
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
bench(lambda: nt(A, B, C), 2.0*M*N*K, f"NT {M}x{N}x{K} (conv1-shaped)")
torch.backends.cuda.matmul.allow_tf32 = True
bench(lambda: torch.matmul(A, B.t(), out=C), 2.0*M*N*K, "cuBLAS tf32 same shape")
M, N, K = 250000, 512, 512
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
bench(lambda: nt(A, B, C), 2.0*M*N*K, f"NT {M}x{N}x{K} (generator hidden)")
bench(lambda: torch.matmul(A, B.t(), out=C), 2.0*M*N*K, "cuBLAS tf32 same shape")
R, Ma, Nb = 250000, 512, 512
P = torch.randn(R, Ma, device="cuda"); Q = torch.randn(R, Nb, device="cuda"); C = torch.zeros(Ma, Nb, device="cuda")
bench(lambda: tn(P, Q, C), 2.0*R*Ma*Nb, f"TN R={R} {Ma}x{Nb} (wgrad)")
bench(lambda: torch.matmul(P.t(), Q, out=C), 2.0*R*Ma*Nb, "cuBLAS tf32 same shape")
