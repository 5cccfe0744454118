#!/usr/bin/env python
"""Where do a large minibatch and its shards of 32 diverge in the backward pass?  Records the per-image outputs of the
generator backward (d_z, d_theta, d_dx) and of the attention backward (d_heads) for both and compares them per image.
python tools/bwd_stage_probe.py cfg5 224"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "target-vae_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
from tvae_b200 import ops, synth
from tvae_b200.config import PRESETS

cfg = PRESETS[sys.argv[1]]
B = int(sys.argv[2])
ctx = bench.Ctx()
wl = bench.Workload(ctx, cfg, 4)
dev = ctx.dev
data = synth.minibatch(cfg, B, seed=9000)
nz = synth.noise(cfg, B, seed=77)
rec = {}
_gb, _ab = ops.generator_bwd, ops.attn_bwd


def gen_bwd(*a, **k):
    out = _gb(*a, **k)
    for key in ("d_z", "d_theta", "d_dx"):
        rec.setdefault(key, []).append(out[key].detach().clone())
    sc = out["scales"].detach().cpu().numpy()
    print("  generator scales: s_i =", [float(sc[2 * i]) for i in range(3)], "colmax", [float(v) for v in sc[20:23]], "amax(d_yhat)", float(sc[31]),
          "| B =", out["d_z"].shape[0], flush=True)
    return out


def attn_bwd(*a, **k):
    d = _ab(*a, **k)
    rec.setdefault("d_heads", []).append(d.detach().flatten(1).clone())
    return d


_ga = ops.gaussian


def gaussian(y_hat, *a, **k):
    out = _ga(y_hat, *a, **k)
    per_image = cfg.n * cfg.n * cfg.n_out
    if out[1] is not None:                       # the backward call: d_yhat
        rec.setdefault("d_yhat", []).append(out[1].detach().reshape(-1, per_image).clone())
    else:
        rec.setdefault("y_hat", []).append(y_hat.detach().reshape(-1, per_image).clone())
    return out


ops.generator_bwd, ops.attn_bwd, ops.gaussian = gen_bwd, attn_bwd, gaussian


def run(lo, hi):
    y = torch.from_numpy(data["y"][lo:hi]).to(dev)
    c = None if data["ctf"] is None else torch.from_numpy(data["ctf"][lo:hi]).to(dev)
    noise = {k: torch.from_numpy(v[lo:hi].copy()).to(dev) for k, v in nz.items()}
    wl.step(y, c, noise=noise, sync=None)
    torch.cuda.synchronize()


run(0, B)
full = {k: v[0] for k, v in rec.items()}
rec.clear()
for lo in range(0, B, 32):
    run(lo, lo + 32)
# the shard steps carry 1 / 32 instead of 1 / B in their loss weights
for k in full:
    sh = torch.cat(rec[k]) * (1.0 if k == "y_hat" else 32.0 / B)
    a, b = full[k].flatten(1) if full[k].dim() > 1 else full[k].reshape(B, 1), sh.flatten(1) if sh.dim() > 1 else sh.reshape(B, 1)
    e = ((a - b).norm(dim=1) / (b.norm(dim=1) + 1e-30)).cpu().numpy()
    bad = np.nonzero(e > 1e-4)[0]
    print(f"{k:8s}: per-image rel err max {e.max():.3e} (image {e.argmax()}), median {np.median(e):.2e}; images > 1e-4: {len(bad)} {bad.tolist()[:24]}")
