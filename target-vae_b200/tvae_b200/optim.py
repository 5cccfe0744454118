"""Fused optimiser step for the training loop.

`Adam` is a drop-in for the reference's `torch.optim.Adam(params, lr=lr)` (train_mnist.py:579): same constructor
arguments, same `param_groups` (so `ReduceLROnPlateau(optim, ...)`, train_mnist.py:581, keeps working), same
`state` keys (`step`, `exp_avg`, `exp_avg_sq`: `state_dict()` is interchangeable with torch's), but `step()` updates
every parameter tensor of a group with ONE launch of the multi-tensor kernel behind `tvae_adam_step`.
There is no CPU path: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes

import torch

from . import ops
from ._lib import check, stream_ptr


class AdamTensor(ctypes.Structure):
    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p),
                ("exp_avg_sq", ctypes.c_void_p), ("numel", ctypes.c_longlong)]


def _lib():
    lib = ops.L()
    if not getattr(lib, "_tvae_optim_configured", False):
        lib.tvae_adam_step.restype = ctypes.c_int
        lib.tvae_adam_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.tvae_running_means.restype = ctypes.c_int
        lib.tvae_running_means.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                                           ctypes.c_void_p, ctypes.c_void_p]
        lib._tvae_optim_configured = True
    return lib


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (amsgrad / maximize / capturable are not part of the reference's use)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _update(self, group, items, zero_grad):
        """One multi-tensor launch per distinct step count over items = [(param, grad)] of `group`."""
        lib = _lib()
        beta1, beta2 = group["betas"]
        todo = []
        for p, g in items:
            if not p.is_cuda:
                raise RuntimeError("tvae_b200.optim.Adam: parameters must be CUDA tensors (no CPU fallback)")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("tvae_b200.optim.Adam: parameters must be contiguous fp32")
            st = self.state[p]
            if not st:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] += 1
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise RuntimeError("tvae_b200.optim.Adam: gradients must be contiguous fp32")
            todo.append((int(st["step"]), p, g, st["exp_avg"], st["exp_avg_sq"]))
        # parameters of a group normally share the step count; launch once per distinct count otherwise
        for k in sorted({t[0] for t in todo}):
            sel = [t for t in todo if t[0] == k]
            table = (AdamTensor * len(sel))()
            for i, (_, p, g, m, v) in enumerate(sel):
                table[i] = AdamTensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
            dev = sel[0][1].device
            if any(t[1].device != dev for t in sel):
                raise RuntimeError("tvae_b200.optim.Adam: the parameters of one group must live on one CUDA device")
            with torch.cuda.device(dev):      # the library launches on the current device
                check(lib.tvae_adam_step(ctypes.cast(table, ctypes.c_void_p), len(sel), float(group["lr"]), float(beta1),
                                         float(beta2), float(group["eps"]), float(group["weight_decay"]), k,
                                         1 if zero_grad else 0, stream_ptr()), "tvae_adam_step")

    @torch.no_grad()
    def step(self, closure=None, zero_grad=False):
        """One update of every parameter that has a gradient; `zero_grad=True` also clears the gradients in the same
        launch (the reference's `optim.step(); optim.zero_grad()` pair, with set_to_none=False semantics)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            items = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    p.grad = g
                items.append((p, g))
            self._update(group, items, zero_grad)
        return loss

    @torch.no_grad()
    def step_tensors(self, params, grads):
        """The same update for an explicit list of parameters with explicit gradient tensors (not `p.grad`), on the current
        stream: the epilogue of a data-parallel gradient bucket (tvae_b200.dp.GradSync(optimizer=...)) - the all-reduced
        bucket is consumed in place, one launch per bucket, while the rest of the backward pass is still running."""
        where = getattr(self, "_group_of", None)
        if where is None or len(where) != sum(len(g["params"]) for g in self.param_groups):
            where = self._group_of = {id(p): g for g in self.param_groups for p in g["params"]}
        by_group = {}
        for p, g in zip(params, grads):
            grp = where.get(id(p))
            if grp is None:
                raise ValueError("step_tensors: parameter does not belong to this optimizer")
            by_group.setdefault(id(grp), (grp, []))[1].append((p, g))
        for grp, items in by_group.values():
            self._update(grp, items, False)


class RunningMeans:
    """ELBO / error / KL running means of train_mnist.py:326-338 kept on the device: `update` enqueues one tiny kernel,
    `read()` is the only host synchronisation (once per epoch or logging interval)."""

    def __init__(self, device):
        self.state = torch.zeros(4, device=device, dtype=torch.float32)

    def update(self, elbo, log_p, kl, b):
        e, l, k = (ops.f32(t).reshape(1) for t in (elbo, log_p, kl))
        with torch.cuda.device(self.state.device):
            check(_lib().tvae_running_means(e.data_ptr(), l.data_ptr(), k.data_ptr(), float(b), self.state.data_ptr(),
                                            stream_ptr()), "tvae_running_means")

    def read(self):
        c, elbo, err, kl = self.state.tolist()
        return elbo, err, kl
