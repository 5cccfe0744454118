"""One training minibatch (eval_minibatch forward + (-elbo).backward()) of a fixed shape as ONE CUDA graph.

The fused step enqueues 35-60 kernels (ours + torch's noise draws and scalar arithmetic) and never synchronises with the
host, so the whole pass is capturable; a replay removes the gaps between them (measured: cfg1 2.74 -> 2.55 ms/step, cfg2
6.48 -> 6.46, the large configs unchanged - the eager step is already 99 % kernel time).  Same call convention as the trainers'
`eval_minibatch` (train_mnist.py:26, train_particles.py:28) - built once per (shape, models), then called per minibatch:

    step = GraphedStep(x_coord, y_shape, generator_model, encoder_model, t_inf, r_inf, device, theta_prior, groupconv,
                       image_dim)                      # particles: ctf_shape=..., padding=..., mask_radius=...
    elbo, log_p_x_g_z, kl_div = step(y)                # every parameter's .grad holds the gradients of -elbo afterwards
    optim.step()

What is static: the input staging buffers (`y`, `ctf`: the call copies the minibatch into them - a host tensor goes
host -> device there, pinned memory makes it asynchronous), the three 0-d results and every parameter's `.grad`
(allocated from the graph's private pool during capture; a replay overwrites them, so `optim.zero_grad(set_to_none=True)`
must NOT be used between replays - `GraphedStep` keeps and re-attaches the tensors).  The noise (Gumbel / normal draws,
models.py:387, train_mnist.py:206,230) comes from torch's CUDA generator inside the graph; torch advances the Philox
offset per replay, so every replay draws fresh noise exactly like the eager call.

The optimiser step stays outside the graph (its bias-correction step count is a host scalar of `tvae_adam_step`).
Data-parallel runs pass `sync=dp.GradSync()` (without a fused optimiser): the two bucket all-reduces are NCCL launches on
NCCL's stream, forked from and joined back into the capturing stream, so they become nodes of the same graph and still
overlap the encoder backward.  Drop the GraphedStep before `destroy_process_group()`: a live graph that holds captured NCCL
launches stalls the communicator's teardown (measured: two minutes).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import elbo as E
from . import ops


class GraphedStep:
    def __init__(self, x_coord, y_shape: Sequence[int], generator_model, encoder_model, t_inf, r_inf, device, theta_prior,
                 groupconv, image_dim=None, *, ctf_shape: Optional[Sequence[int]] = None, particles: bool = False, padding: int = 0,
                 mask_radius: int = 0, sync=None, warmup: int = 3):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("GraphedStep: the hot path runs on sm_100a only (no CPU fallback)")
        self.device = device
        self.particles = bool(particles or ctf_shape is not None)
        self.x = x_coord
        self.gen, self.enc = generator_model, encoder_model
        self.params = [p for p in list(generator_model.parameters()) + list(encoder_model.parameters()) if p.requires_grad]
        self._args = (t_inf, r_inf, theta_prior, groupconv, image_dim, padding, mask_radius)
        if sync is not None and getattr(sync, "optimizer", None) is not None:
            raise ValueError("GraphedStep: a GradSync with a fused optimiser cannot be captured (the Adam step count is a host "
                             "scalar); pass dp.GradSync() and call optim.step() after the replay")
        self.sync = sync
        with torch.cuda.device(device):
            self.y = torch.zeros(*y_shape, device=device, dtype=torch.float32)
            self.ctf = None if ctf_shape is None else torch.zeros(*ctf_shape, device=device, dtype=torch.float32)
            if self.ctf is not None:
                # a valid filter for the warm-up passes: the identity (a centred delta)
                m = self.ctf.shape[-1]
                self.ctf.reshape(self.ctf.shape[0], m, m)[:, m // 2, m // 2] = 1.0
            # warm-up on a side stream (torch's capture recipe): fills every host-side cache of the path - TMA descriptors,
            # function attributes, the pixel spacing (its one host read), log-prior / head tables, sigma scalar
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self._eager()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            for p in self.params:
                p.grad = None
            was_profiling = ops.profile_enabled()
            ops.profile_enable(False)          # per-kernel CUDA events cannot be read back from a captured stream
            self.graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            try:
                # thread_local: only THIS thread's calls are checked against the capture.  Under the default ("global") a CUDA
                # call of any other thread - ProcessGroupNCCL's watchdog polling the events of earlier collectives - invalidates
                # it (seen with two ranks: cudaErrorStreamCaptureInvalidated); the step itself makes no unsafe call on any thread
                with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                    elbo, logp, kl, grads = self._eager()
            finally:
                ops.profile_enable(was_profiling)
            self.launches_per_replay = ops.launch_count() - n0      # library launches inside one replay
            self.elbo, self.log_p_x_g_z, self.kl_div = elbo.detach(), logp.detach(), kl.detach()
            self.grads = [None if g is None else g.detach() for g in grads]
            for p, g in zip(self.params, self.grads):
                p.grad = g

    def _eager(self):
        """One pass.  The step runs on fresh leaf ALIASES of the parameters (same storage) and takes its gradients with
        torch.autograd.grad: no AccumulateGrad node of the real parameters takes part.  Those nodes are cached on the parameter
        while any earlier graph of the caller is alive (a kept `elbo`), and they are bound to the stream they were created on -
        usually the legacy default stream, which a capturing stream must never be joined into
        (cudaErrorStreamCaptureImplicit).  The gradients are what the fused node returns: views of its flat buckets."""
        t_inf, r_inf, theta_prior, groupconv, image_dim, padding, mask_radius = self._args
        aliases = {}

        def alias(t):
            if not t.requires_grad:
                return t
            a = aliases.get(id(t))
            if a is None:
                a = aliases[id(t)] = t.detach().requires_grad_(True)
            return a
        if self.particles:
            elbo, logp, kl = E.eval_minibatch_particles(self.x, self.y, self.ctf, self.gen, self.enc, t_inf, r_inf, 0, self.device,
                                                        theta_prior, groupconv, padding, mask_radius, sync=self.sync, _alias=alias)
        else:
            elbo, logp, kl = E.eval_minibatch(self.x, self.y, self.gen, self.enc, t_inf, r_inf, 0, self.device, theta_prior,
                                              groupconv, image_dim, sync=self.sync, _alias=alias)
        used = [p for p in self.params if id(p) in aliases]
        got = torch.autograd.grad(-elbo, [aliases[id(p)] for p in used], allow_unused=True)
        by_id = {id(p): g for p, g in zip(used, got)}
        return elbo, logp, kl, [by_id.get(id(p)) for p in self.params]

    def __call__(self, y, ctf=None):
        """Copies the minibatch into the staging buffers (any device; pinned host memory copies asynchronously), replays
        the step.  -> (elbo, log_p_x_g_z, kl_div): static 0-d device tensors, overwritten by the next call."""
        if tuple(y.shape) != tuple(self.y.shape):
            raise ValueError(f"GraphedStep: captured for y {tuple(self.y.shape)}, got {tuple(y.shape)}; build another GraphedStep "
                             "for the last, shorter minibatch of an epoch (or run it through eval_minibatch)")
        if (ctf is None) != (self.ctf is None):
            raise ValueError("GraphedStep: ctf must be passed exactly when the step was captured with one")
        self.y.copy_(y, non_blocking=True)
        if ctf is not None:
            if ctf.numel() != self.ctf.numel():
                raise ValueError(f"GraphedStep: captured for ctf {tuple(self.ctf.shape)}, got {tuple(ctf.shape)}")
            self.ctf.copy_(ctf.reshape(self.ctf.shape), non_blocking=True)
        for p, g in zip(self.params, self.grads):
            if p.grad is not g:
                p.grad = g                      # re-attach after an optimiser's zero_grad(set_to_none=True) / an eager backward
        self.graph.replay()
        return self.elbo, self.log_p_x_g_z, self.kl_div
