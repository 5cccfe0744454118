#!/usr/bin/env python
"""N eager fwd+bwd steps of one config and nothing else (for ncu launch lists): python tools/eager_steps.py cfg1 [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "target-vae_b200")):
    sys.path.insert(0, p)
import torch
import bench
from tvae_b200.config import PRESETS
cfg = PRESETS[sys.argv[1]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = bench.Workload(bench.Ctx(), cfg, cfg.batch)
for i in range(n):
    wl.step_resident(i)
torch.cuda.synchronize()
