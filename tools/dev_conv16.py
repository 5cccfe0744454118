#!/usr/bin/env python
"""Dev check of the fp16 group-conv kernels against torch (GPU box only): forward and weight gradient w.r.t. the bank."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "target-vae_b200")]
import torch
import torch.nn.functional as F
from tvae_b200 import ops
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"

def run(B, C, n, k, p, G, O, timing=False):
    torch.manual_seed(0)
    s = ops.enc_shape(B, C, n, k, p, G, O, 2)
    d = n + 2 * p - k + 1
    w = (torch.rand(O, C, 1, k, k, device=dev) * 2 - 1) / (C * k * k) ** 0.5
    y = torch.rand(B, C, n, n, device=dev)
    bias = torch.randn(O, device=dev) * 0.1
    bank = ops.filter_bank_fwd(s, w)
    K = C * k * k
    out = ops.empty(B * G * d * d, O, device=dev)
    L = ops.L()
    ops.check(L.tvae_groupconv_fwd(ops.byref(s), ops.ptr(y), ops.ptr(bank), ops.ptr(bias), ops.ptr(out), ops.stream_ptr()), "fwd")
    torch.cuda.synchronize()
    tw = bank[:, :K].float().view(G * O, C, k, k)      # rows r*O + o
    ref = F.conv2d(y, tw, None, 1, p).view(B, G, O, d, d).permute(0, 1, 3, 4, 2) + bias
    ref = ref.reshape(-1, O)
    e_f = float((out - ref).norm() / ref.norm())
    # wgrad
    g = torch.randn(B * G * d * d, O, device=dev) * 1e-3
    g16 = torch.empty_like(g, dtype=torch.float16)
    scales = ops.empty(8, device=dev)
    dbank = ops.empty(G * O, s.kpad, device=dev)
    ops.check(L.tvae_groupconv_wgrad(ops.byref(s), ops.ptr(y), ops.ptr(g), ops.ptr(g16), ops.ptr(scales), ops.ptr(dbank), ops.stream_ptr()), "wgrad")
    torch.cuda.synchronize()
    twr = tw.clone().requires_grad_(True)
    o2 = F.conv2d(y, twr, None, 1, p)                   # (B, G*O, d, d)
    gg = g.view(B, G, d, d, O).permute(0, 1, 4, 2, 3).reshape(B, G * O, d, d)
    (o2 * gg).sum().backward()
    refw = twr.grad.view(G * O, K)
    e_w = float((dbank[:, :K] - refw).norm() / refw.norm())
    refb = g.view(B * G, d * d, O).sum((0, 1))
    gotb = dbank[:O, K]
    e_b = float((gotb - refb).norm() / refb.norm())
    e_16 = float((g16.float() * float(scales[3]) - g).norm() / g.norm())
    msg = f"B{B} C{C} n{n} k{k} p{p} G{G} O{O}: fwd {e_f:.2e} wgrad {e_w:.2e} bias {e_b:.2e} (fp16 copy {e_16:.2e}, scale {float(scales[2]):.3g})"
    if timing:
        for fn, nm in ((lambda: L.tvae_groupconv_fwd(ops.byref(s), ops.ptr(y), ops.ptr(bank), ops.ptr(bias), ops.ptr(out), ops.stream_ptr()), "fwd"),
                       (lambda: L.tvae_groupconv_wgrad(ops.byref(s), ops.ptr(y), ops.ptr(g), ops.ptr(g16), ops.ptr(scales), ops.ptr(dbank), ops.stream_ptr()), "wgrad(+cvt)")):
            for _ in range(3): fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            msg += f" | {nm} {e0.elapsed_time(e1) / 10:.3f} ms"
    print(msg, flush=True)

if __name__ == "__main__":
    if "--one" in sys.argv:
        run(2, 1, 16, 16, 8, 4, 32)
        sys.exit(0)
    run(2, 1, 16, 16, 8, 4, 32)
    run(3, 1, 20, 9, 3, 8, 32)
    run(2, 3, 12, 12, 6, 8, 32)
    run(3, 1, 16, 9, 2, 16, 32)
    run(2, 1, 28, 12, 4, 8, 64)
    run(4, 1, 50, 28, 8, 8, 128)
    run(4, 1, 64, 64, 32, 8, 128)
    run(2, 3, 64, 64, 32, 8, 128)
    run(2, 1, 128, 64, 16, 16, 128)
    if "--time" in sys.argv:
        run(100, 1, 64, 64, 32, 8, 128, timing=True)
        run(100, 1, 50, 28, 8, 8, 128, timing=True)
