"""Data-parallel sharding of the hot path: one process per GPU, each rank takes a contiguous
slice of the global minibatch, holds a full replica and all-reduces (averages) the gradients over NCCL/NVLink.

The ELBO is a mean over images of per-image terms (train_mnist.py:282,291), so with equal shards the average of
the per-rank gradients equals the single-GPU large-batch gradient.  The only exchange step is the gradient
all-reduce: 0.8-2.9 M fp32 values (3-12 MB), latency-bound on NVSwitch.  It is issued as two buckets - the
generator's as soon as the generator backward kernels have been enqueued, so that it runs on NCCL's stream
underneath the encoder backward; the encoder's at the end - and nothing else crosses ranks on the data path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of rank `rank`; requires equal shards (averaging identity)."""
    if global_batch % world != 0:
        raise ValueError(f"global minibatch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard(t: Optional[torch.Tensor], rank: int, world: int):
    if t is None:
        return None
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi].contiguous()


class GradSync:
    """Two-bucket gradient averaging, in place, with an optional fused optimiser epilogue.

    The backward kernels write each bucket's parameter gradients straight into ONE flat fp32 buffer (ops.flat_views);
    start(i, flat) launches an async all-reduce on that buffer - no flatten copy - and finish() makes the current stream
    wait for both (a stream-side wait, the host never blocks).  The per-parameter gradients handed to autograd are views
    of the buckets, so they hold the averaged values afterwards.

    optimizer = tvae_b200.optim.Adam: SURVEY.md 8f-1 "optimiser step fused with the all-reduce epilogue" - as soon as a
    bucket's collective has finished (on a side stream: the backward pass keeps running), ONE multi-tensor Adam launch
    updates that bucket's parameters from the averaged bucket (`optim.step()` must then not be called again for the step;
    tvae_b200.train.train_epoch does this).  The generator's update runs underneath the encoder backward, the encoder's
    right behind its all-reduce.  Works for a single process too (no collective, same epilogue)."""

    def __init__(self, group=None, optimizer=None):
        self.group = group
        self.optimizer = optimizer
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._pending: List = [None, None]
        self._side = {}
        # NCCL averages inside the collective (no separate scaling kernel); gloo (CPU tests) only sums
        self._avg = dist.is_initialized() and self.world > 1 and dist.get_backend(group) == "nccl"

    def _side_stream(self, device):
        s = self._side.get(device)
        if s is None:
            s = self._side[device] = torch.cuda.Stream(device=device)
        return s

    def start(self, bucket: int, flat: torch.Tensor, params: Optional[Sequence[torch.Tensor]] = None,
              grads: Optional[Sequence[torch.Tensor]] = None):
        """flat: the bucket; params / grads (views of `flat`, one per parameter): what the fused optimiser epilogue updates."""
        work = None
        if self.world > 1:
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM, group=self.group, async_op=True)
        side = None
        if self.optimizer is not None and params is not None and flat.is_cuda:
            side = self._side_stream(flat.device)
            side.wait_stream(torch.cuda.current_stream(flat.device))       # the kernels that filled the bucket
            with torch.cuda.stream(side):
                if work is not None:
                    work.wait()                                            # the SIDE stream waits for the collective
                    if not self._avg:
                        flat.mul_(1.0 / self.world)
                    work = None
                self.optimizer.step_tensors(params, grads)
                flat.record_stream(side)
        self._pending[bucket] = (flat, work, side)

    def _resolve(self, bucket: int):
        flat, work, side = self._pending[bucket]
        if work is not None:
            work.wait()
            if not self._avg:
                flat.mul_(1.0 / self.world)
        if side is not None:
            torch.cuda.current_stream(flat.device).wait_stream(side)        # parameters are updated before anything that follows
        self._pending[bucket] = None
        return flat

    def finish(self):
        return self._resolve(0), self._resolve(1)


def all_reduce_scalars(vals: torch.Tensor, group=None) -> torch.Tensor:
    """Average logged scalars (elbo / error / kl) across ranks, once per logging interval."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vals, group=group)
        vals = vals / dist.get_world_size(group)
    return vals
