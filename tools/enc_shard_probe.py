#!/usr/bin/env python
"""Encoder forward alone: head maps of one large batch against shards of 32 images, per image.
python tools/enc_shard_probe.py cfg5 224"""
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
B = int(sys.argv[2])
dev = torch.device("cuda", 0)
gen, enc = bench.build_models(cfg, dev)
y = torch.from_numpy(synth.minibatch(cfg, B, seed=9000)["y"]).to(dev)
with torch.no_grad():
    full = enc.head_maps(y).flatten(1)
    worst = []
    for lo in range(0, B, 32):
        part = enc.head_maps(y[lo:lo + 32].contiguous()).flatten(1)
        d = (full[lo:lo + 32] - part).norm(dim=1) / part.norm(dim=1)
        worst += d.cpu().tolist()
worst = np.array(worst)
print(f"{cfg.name} encoder forward alone, B = {B}: per-image head-map error max {worst.max():.3e} (image {worst.argmax()}); images > 1e-6: {np.nonzero(worst > 1e-6)[0].tolist()[:50]}")
