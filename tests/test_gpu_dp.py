"""Data-parallel hot path on real GPUs: 2 ranks over NCCL, each on a contiguous half of the global
minibatch with the two-bucket GradSync issued from inside the backward, must reproduce the single-GPU full-batch
ELBO and gradients (equal shards => averaged gradients == large-batch gradient up to fp32 summation order).
Skipped when fewer than 2 GPUs are visible (the CPU/gloo version of the identity is tests/test_dp_gloo.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import test_gpu_step as T
        from tvae_b200 import dp, elbo as E, synth
        from tvae_b200.config import CFG1
        T.DEV = dev
        cfg, B = CFG1.with_(name="cfg1_dp"), 16
        gen, enc = T.build_models(cfg)
        data = synth.minibatch(cfg, B, 0)
        nz = {k: torch.from_numpy(v).to(dev) for k, v in synth.noise(cfg, B, 0).items()}
        x = torch.from_numpy(synth.image_coords(cfg.n)).to(dev)
        y = torch.from_numpy(data["y"]).to(dev)
        params = list(enc.named_parameters()) + list(gen.named_parameters())
        # single-GPU full batch (every rank computes it locally; no communication)
        e_full, _, _ = E.eval_minibatch(x, y, gen, enc, "attention", "attention+offsets", 0, dev, cfg.theta_prior, cfg.G, cfg.n, noise=nz)
        (-e_full).backward()
        g_full = {k: p.grad.clone() for k, p in params}
        for _, p in params:
            p.grad = None
        # sharded step with gradient averaging
        sync = dp.GradSync()
        ys = dp.shard(y, rank, world)
        nzs = {k: dp.shard(v, rank, world) for k, v in nz.items()}
        e, _, _ = E.eval_minibatch(x, ys, gen, enc, "attention", "attention+offsets", 0, dev, cfg.theta_prior, cfg.G, cfg.n,
                                   noise=nzs, sync=sync)
        (-e).backward()
        e_mean = dp.all_reduce_scalars(e.detach().clone().reshape(1))
        torch.cuda.synchronize()
        worst = 0.0
        for k, p in params:
            if k == "conv_a.bias":
                continue
            err = float((p.grad - g_full[k]).norm() / (g_full[k].norm() + 1e-30))
            worst = max(worst, err)
        ok = worst < 2e-3 and abs(float(e_mean) - float(e_full)) < 1e-5 * abs(float(e_full))
        # the same sharded step as ONE CUDA graph (the two NCCL bucket all-reduces are nodes of it): replay == eager call on the
        # same generator seed (noise drawn inside, different per rank)
        from tvae_b200.graph import GraphedStep
        gs = GraphedStep(x, ys.shape, gen, enc, "attention", "attention+offsets", dev, cfg.theta_prior, cfg.G, cfg.n, sync=dp.GradSync())
        for _, p in params:
            p.grad = None
        torch.manual_seed(500 + rank)
        e2, _, _ = E.eval_minibatch(x, ys, gen, enc, "attention", "attention+offsets", 0, dev, cfg.theta_prior, cfg.G, cfg.n, sync=sync)
        (-e2).backward()
        torch.cuda.synchronize()
        g_eager = {k: p.grad.clone() for k, p in params}
        torch.manual_seed(500 + rank)
        e3 = float(gs(ys)[0])
        worst_g = max(float((p.grad - g_eager[k]).norm() / (g_eager[k].norm() + 1e-30)) for k, p in params if k != "conv_a.bias")
        ok = ok and worst_g < 5e-4 and abs(e3 - float(e2)) < 1e-5 * abs(float(e2))
        del gs                               # a live graph holding captured NCCL launches stalls destroy_process_group
        torch.cuda.synchronize()
        q.put((rank, bool(ok), worst, float(e_mean), float(e_full), worst_g))
    except Exception as exc:                 # report instead of leaving the parent to time out on the queue
        import traceback
        q.put((rank, False, f"{type(exc).__name__}: {exc}", traceback.format_exc()[-1500:]))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_matches_full_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    print(res)
    assert all(r[1] for r in res), res
