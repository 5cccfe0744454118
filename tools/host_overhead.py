#!/usr/bin/env python
"""Host-side cost of one fwd+bwd step (enqueue time without synchronising) against its device time:
    python tools/host_overhead.py cfg1 [B]
A config whose enqueue time exceeds the device time is launch-bound."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "target-vae_b200")):
    sys.path.insert(0, p)
import torch
import bench
from tvae_b200.config import PRESETS

cfg = PRESETS[sys.argv[1] if len(sys.argv) > 1 else "cfg1"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch
ctx = bench.Ctx()
wl = bench.Workload(ctx, cfg, B)
for i in range(5):
    wl.step_resident(i)
torch.cuda.synchronize()
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(K):
    wl.step_resident(i)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{cfg.name} B={B}: host enqueue {1e3 * (t1 - t0) / K:.3f} ms/step, device {e0.elapsed_time(e1) / K:.3f} ms/step, "
      f"wall {1e3 * (t2 - t0) / K:.3f} ms/step")
if "--profile" in sys.argv:
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(K):
        wl.step_resident(i)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
