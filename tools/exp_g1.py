"""Experiment: does the kernel path handle groupconv = 1 (no rotation)?  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "target-vae_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from helpers import oracle_step, rel_err
from tvae_b200.config import HotPathConfig
import test_gpu_step as T

for cfg, B in ((HotPathConfig("g1a", C=1, n=20, k=20, p=10, G=1, z=2, O=32, hidden=64, rot_refinement=False), 3),
               (HotPathConfig("g1b", C=1, n=28, k=28, p=14, G=1, z=2, O=128, hidden=128, rot_refinement=False), 4)):
    elbo, logp, kl, grads = T.run_step(cfg, B)
    o_elbo, o_logp, o_kl, _, o_grads = oracle_step(cfg, B, dtype=torch.float64)
    print(cfg.name, elbo, float(o_elbo), logp, float(o_logp), kl, float(o_kl))
    for k, v in grads.items():
        print("   ", k, f"{rel_err(v, o_grads[k]):.2e}")
