#!/usr/bin/env python
"""Gradient of one large minibatch against the mean of the gradients of its equal shards, on ONE GPU (the data-parallel
identity of SURVEY 8e without any communication): python tools/shard_consistency.py cfg5 256 8
prints the relative Frobenius error per parameter.  Same images, same injected noise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "target-vae_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
from tvae_b200 import synth
from tvae_b200.config import PRESETS

cfg = PRESETS[sys.argv[1]]
B, S = int(sys.argv[2]), int(sys.argv[3])
for kv in sys.argv[4:]:                      # overrides, e.g. ctf=0 z=2
    k, v = kv.split("=")
    cfg = cfg.with_(**{k: type(getattr(cfg, k))(int(v))})
ctx = bench.Ctx()
wl = bench.Workload(ctx, cfg, 4)
dev = ctx.dev
data = synth.minibatch(cfg, B, seed=9000)
nz = synth.noise(cfg, B, seed=77)
names = [n for n, _ in wl.gen.named_parameters()] + [n for n, _ in wl.enc.named_parameters()]


def grads(lo, hi):
    y = torch.from_numpy(data["y"][lo:hi]).to(dev)
    c = None if data["ctf"] is None else torch.from_numpy(data["ctf"][lo:hi]).to(dev)
    noise = {k: torch.from_numpy(v[lo:hi].copy()).to(dev) for k, v in nz.items()}
    e = wl.step(y, c, noise=noise, sync=None)
    torch.cuda.synchronize()
    return float(e), [p.grad.detach().double().clone() for p in wl.params]


e_full, g_full = grads(0, B)
per = B // S
acc, es = None, []
for s in range(S):
    e, g = grads(s * per, (s + 1) * per)
    es.append(e)
    acc = g if acc is None else [a + b for a, b in zip(acc, g)]
g_mean = [a / S for a in acc]
print(f"{cfg.name}: B = {B} against {S} shards of {per}; ELBO full {e_full:.4f}, mean of shards {np.mean(es):.4f}")
for n, a, b in zip(names, g_full, g_mean):
    print(f"  {n:28s} |grad| {float(a.norm()):10.4g}   rel err {float((a - b).norm() / (a.norm() + 1e-30)):.3e}")
