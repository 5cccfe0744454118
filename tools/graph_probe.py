#!/usr/bin/env python
"""Eager step against the CUDA-graph replay of the same step (tvae_b200.graph.GraphedStep):
    python tools/graph_probe.py cfg1 cfg2 cfg4
prints device ms/step for both (inputs resident in HBM) and the ELBO / gradient agreement on identical seeds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "target-vae_b200")):
    sys.path.insert(0, p)
import torch
import bench
from tvae_b200.config import PRESETS
from tvae_b200.graph import GraphedStep

ctx = bench.Ctx()
for name in (sys.argv[1:] or ["cfg1", "cfg2"]):
    cfg = PRESETS[name]
    B = cfg.batch
    wl = bench.Workload(ctx, cfg, B)
    K = 20 if name != "cfg5" else 4
    for i in range(4):
        wl.step_resident(i)
    ms_eager = ctx.timed(wl.step_resident, K) / K
    gs = wl.graphed()
    for i in range(3):
        gs(wl.y_dev[i % wl.NB], wl.ctf_dev[i % wl.NB])
    ms_graph = ctx.timed(lambda i: gs(wl.y_dev[i % wl.NB], wl.ctf_dev[i % wl.NB]), K) / K
    ms_eager2 = ctx.timed(wl.step_resident, K) / K
    ms_graph2 = ctx.timed(lambda i: gs(wl.y_dev[i % wl.NB], wl.ctf_dev[i % wl.NB]), K) / K

    def synced(i):
        gs(wl.y_dev[i % wl.NB], wl.ctf_dev[i % wl.NB])
        torch.cuda.synchronize()
    ms_graph_sync = ctx.timed(synced, K) / K
    if ctx.rank == 0:
        print(f"{cfg.name}: second pass eager {ms_eager2:.3f}, graph {ms_graph2:.3f}, graph with a host sync per step {ms_graph_sync:.3f} ms/step", flush=True)
    # agreement on identical seeds
    torch.manual_seed(1234)
    e0 = wl.step_resident(1).detach().clone()
    g0 = [p.grad.detach().clone() for p in gs.params]
    torch.manual_seed(1234)
    e1 = gs(wl.y_dev[1], wl.ctf_dev[1])[0].clone()
    torch.cuda.synchronize()
    err = max(float((a - p.grad).norm() / (a.norm() + 1e-30)) for a, p in zip(g0, gs.params) if float(a.norm()) > 1e-12)
    if ctx.rank == 0:
      print(f"{cfg.name} B={B}: eager {ms_eager:.3f} ms/step, graph {ms_graph:.3f} ms/step ({gs.launches_per_replay} library launches per replay); "
          f"elbo eager {float(e0):.6f} graph {float(e1):.6f}, worst relative gradient difference {err:.2e}", flush=True)
    del wl, gs
    torch.cuda.empty_cache()
