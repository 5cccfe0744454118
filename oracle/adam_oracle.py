"""TEST INFRASTRUCTURE ONLY - CPU restatement of the optimiser step and running statistics of the reference's training loop.

The reference calls `torch.optim.Adam(params, lr=lr)` (train_mnist.py:579) then `optim.step(); optim.zero_grad()`
(train_mnist.py:323-324) and keeps running means of ELBO / error / KL with three `.item()` syncs per step
(train_mnist.py:326-338).  The optimiser arithmetic lives in a third-party dependency that is not vendored in the
reference checkout: PyTorch (README pins only ">= 1.11"; the installed 2.11.0 is the de-facto pinned version),
torch/optim/adam.py::_single_tensor_adam with amsgrad=False, maximize=False, foreach=False.  Restated here in numpy
fp32; pinned by tests/test_optim_oracle.py against torch.optim.Adam itself run on the CPU (step-by-step, several
steps, with and without weight decay).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr=2e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """One torch.optim.Adam update (step is 1-based).  All arrays fp32; returns new (param, exp_avg, exp_avg_sq)."""
    f = np.float32
    p, g, m, v = (np.asarray(a, dtype=f) for a in (param, grad, exp_avg, exp_avg_sq))
    if weight_decay != 0:
        g = g + f(weight_decay) * p                                  # grad.add(param, alpha=weight_decay)
    m = m + (g - m) * f(1 - beta1)                                   # exp_avg.lerp_(grad, 1 - beta1)
    v = v * f(beta2) + f(1 - beta2) * g * g                          # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bias_correction1 = 1 - beta1 ** step                             # python floats (double), adam.py
    bias_correction2 = 1 - beta2 ** step
    step_size = lr / bias_correction1
    bias_correction2_sqrt = math.sqrt(bias_correction2)
    denom = np.sqrt(v) / f(bias_correction2_sqrt) + f(eps)           # (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p - f(step_size) * (m / denom)                               # param.addcdiv_(exp_avg, denom, value=-step_size)
    return p.astype(f), m.astype(f), v.astype(f)


def running_means(state, elbo, log_p, kl, b):
    """train_mnist.py:326-338: state = [c, elbo_accum, gen_loss_accum, kl_loss_accum] -> updated copy (fp32)."""
    f = np.float32
    c = f(state[0]) + f(b)
    out = [c]
    for acc, x in zip(state[1:], (elbo, -log_p, kl)):
        delta = f(b) * (f(x) - f(acc))
        out.append(f(acc) + delta / c)
    return np.asarray(out, dtype=f)
