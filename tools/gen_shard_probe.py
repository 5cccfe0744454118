#!/usr/bin/env python
"""SpatialGenerator alone (explicit coordinates): one large batch against its shards of 32 images, per image.
python tools/gen_shard_probe.py cfg5 224"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "target-vae_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
from tvae_b200.config import PRESETS

cfg = PRESETS[sys.argv[1]]
B = int(sys.argv[2])
dev = torch.device("cuda", 0)
gen, enc = bench.build_models(cfg, dev)
g = torch.Generator(device="cpu").manual_seed(5)
N = cfg.n * cfg.n
x = (torch.rand(B, N, 2, generator=g) * 2 - 1).to(dev)
z = torch.randn(B, cfg.z, generator=g).to(dev)
mult = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0      # scales the upstream gradient (moves the device-chosen fp16 scales)
r = torch.randn(B, N, cfg.n_out, generator=g).to(dev) * (mult / B)
params = list(gen.parameters())
names = [n for n, _ in gen.named_parameters()]


def run(lo, hi):
    for p in params:
        p.grad = None
    zz = z[lo:hi].clone().requires_grad_(True)
    y = gen(x[lo:hi].contiguous(), zz)
    (y * r[lo:hi]).sum().backward()
    torch.cuda.synchronize()
    return y.detach().clone(), zz.grad.clone(), [p.grad.detach().double().clone() for p in params]


y_f, dz_f, g_f = run(0, B)
ys, dzs, acc = [], [], None
for lo in range(0, B, 32):
    y, dz, gs = run(lo, min(lo + 32, B))
    ys.append(y); dzs.append(dz)
    acc = gs if acc is None else [a + b for a, b in zip(acc, gs)]
y_s, dz_s = torch.cat(ys), torch.cat(dzs)
ey = ((y_f - y_s).flatten(1).norm(dim=1) / y_s.flatten(1).norm(dim=1)).cpu().numpy()
ez = ((dz_f - dz_s).norm(dim=1) / dz_s.norm(dim=1)).cpu().numpy()
print(f"{cfg.name} generator alone, B = {B}: per-image y_hat error max {ey.max():.2e} (image {ey.argmax()}), d_z error max {ez.max():.2e} (image {ez.argmax()})")
print("  images with d_z error > 1e-3:", np.nonzero(ez > 1e-3)[0].tolist()[:40])
for n, a, b in zip(names, g_f, acc):
    print(f"  {n:28s} rel err {float((a - b).norm() / (b.norm() + 1e-30)):.3e}")
