#!/usr/bin/env python
"""Runs a few fwd+bwd steps of a trainer VARIANT at dSprites size (for ncu captures of the variant-only kernels):
    python tools/run_variant_step.py pooled      --r-inf unimodal --groupconv 8 (rot_pool_fwd / rot_pool_bwd)
    python tools/run_variant_step.py tanh        --activation tanh
Prints CUDA-event ms per step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "target-vae_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from tvae_b200 import elbo as E, synth
from tvae_b200.config import CFG2
import test_gpu_step as T

kind = sys.argv[1] if len(sys.argv) > 1 else "pooled"
B = 100
if kind == "pooled":
    cfg = CFG2.with_(name="cfg2_pooled", rot_refinement=False, normal_prior_over_r=False, encoder="attn_unimodal")
    r_inf = "unimodal"
else:
    cfg = CFG2.with_(name="cfg2_tanh", activation="tanh")
    r_inf = "attention+offsets"
gen, enc = T.build_models(cfg)
x = torch.from_numpy(synth.image_coords(cfg.n)).to("cuda")
y = torch.from_numpy(synth.minibatch(cfg, B, 0)["y"]).to("cuda")
params = list(gen.parameters()) + list(enc.parameters())


def step():
    for p in params:
        p.grad = None
    elbo, _, _ = E.eval_minibatch(x, y, gen, enc, "attention", r_inf, 0, "cuda", cfg.theta_prior, cfg.G, cfg.n)
    (-elbo).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    step()
e1.record()
torch.cuda.synchronize()
print(f"{cfg.name}: {e0.elapsed_time(e1) / K:.3f} ms / step, {B * K / (e0.elapsed_time(e1) * 1e-3):.0f} images/s")
