"""Data-parallel host logic on CPU: world_size-2 gloo run of GradSync + sharding.

Checks the averaging identity the N-GPU path relies on: with equal contiguous shards, the two-bucket averaged
per-rank gradients equal the gradient of the full-batch mean loss.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tvae_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        B, D = 8, 5
        X = torch.randn(B, D, dtype=torch.float64)
        w_gen = torch.randn(D, 3, dtype=torch.float64, requires_grad=True)
        w_enc = torch.randn(3, dtype=torch.float64, requires_grad=True)
        full = ((X @ w_gen).tanh() @ w_enc).pow(2).mean()        # mean over images of per-image terms
        g_full = torch.autograd.grad(full, [w_gen, w_enc])
        xs = dp.shard(X, rank, world)
        assert xs.shape[0] == B // world
        loss = ((xs @ w_gen).tanh() @ w_enc).pow(2).mean()
        g_gen, g_enc = torch.autograd.grad(loss, [w_gen, w_enc])
        # the backward kernels write each bucket's gradients into one flat buffer; the per-parameter gradients are views
        from tvae_b200 import ops
        flat0, (a_gen, a_gen0) = ops.flat_views([tuple(g_gen.shape), tuple(g_gen[0].shape)], "cpu")   # two shapes in one bucket
        flat1, (a_enc,) = ops.flat_views([tuple(g_enc.shape)], "cpu")
        flat0, flat1 = flat0.double(), flat1.double()
        a_gen, a_gen0, a_enc = flat0[:g_gen.numel()].view_as(g_gen), flat0[16:16 + g_gen[0].numel()].view_as(g_gen[0]), flat1[:3]
        a_gen.copy_(g_gen); a_gen0.copy_(g_gen[0]); a_enc.copy_(g_enc)
        sync = dp.GradSync()
        sync.start(0, flat0)
        sync.start(1, flat1)
        r0, r1 = sync.finish()
        ok = (r0 is flat0 and r1 is flat1                                     # averaged IN PLACE: the views see it
              and torch.allclose(a_gen, g_full[0], atol=1e-12) and torch.allclose(a_enc, g_full[1], atol=1e-12)
              and torch.allclose(a_gen0, g_full[0][0], atol=1e-12) and a_gen.shape == g_gen.shape)
        sc = dp.all_reduce_scalars(torch.tensor([float(rank)], dtype=torch.float64))
        ok = ok and abs(float(sc) - (world - 1) / 2) < 1e-12
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gradsync_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds():
    assert dp.shard_bounds(100, 3, 4) == (75, 100)
    with pytest.raises(ValueError):
        dp.shard_bounds(10, 0, 4)
    sync = dp.GradSync()                       # single process: identity, in place
    from tvae_b200 import ops
    flat, (a, b) = ops.flat_views([(2, 3), (5,)], "cpu")
    assert flat.numel() == 8 + 8 and a.data_ptr() == flat.data_ptr() and b.data_ptr() == flat.data_ptr() + 8 * 4   # 16-byte aligned views
    a.copy_(torch.arange(6.0).view(2, 3)); b.fill_(1.0)
    sync.start(0, flat)
    sync.start(1, torch.zeros(4))
    f0, f1 = sync.finish()
    assert f0 is flat and torch.equal(a, torch.arange(6.0).view(2, 3)) and torch.equal(b, torch.ones(5)) and f1.shape == (4,)
