"""CPU: the optimiser oracle (oracle/adam_oracle.py) is pinned against torch.optim.Adam itself - the third-party code
the reference's training loop calls (train_mnist.py:579, 323-324) - and the running-mean restatement against the
reference's literal update (train_mnist.py:326-338).  Host-side validation of the C-ABI entry points needs no GPU."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import adam_oracle as ao


def _close(a, b):
    """fp32 agreement up to re-association: 2e-6 relative, with an absolute floor of 2e-6 x the tensor's scale for
    elements that are themselves the result of a cancellation."""
    np.testing.assert_allclose(a, b, rtol=2e-6, atol=2e-6 * float(np.abs(b).max()) + 1e-30)


@pytest.mark.parametrize("wd", [0.0, 0.01])
@pytest.mark.parametrize("n", [1, 7, 4097])
def test_adam_oracle_matches_torch(n, wd):
    rng = np.random.default_rng(n)
    p0 = rng.standard_normal(n).astype(np.float32)
    p_t = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([p_t], lr=2e-4, weight_decay=wd)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 8):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 2)).astype(np.float32)
        p_t.grad = torch.from_numpy(g.copy())
        opt.step()
        p, m, v = ao.adam_step(p, g, m, v, step, lr=2e-4, weight_decay=wd)
        st = opt.state[p_t]
        _close(m, st["exp_avg"].numpy())
        _close(v, st["exp_avg_sq"].numpy())
        _close(p, p_t.detach().numpy())


def test_running_means_oracle_matches_reference_loop():
    rng = np.random.default_rng(0)
    state = np.zeros(4, np.float32)
    c = 0
    gen_loss_accum = kl_loss_accum = elbo_accum = 0      # train_mnist.py:302-305
    for _ in range(20):
        b = int(rng.integers(1, 101))
        elbo, log_p, kl = float(rng.normal(-500, 50)), float(rng.normal(-480, 50)), float(rng.uniform(1, 30))
        gen_loss, kl_loss = -log_p, kl                     # train_mnist.py:326-338
        c += b
        delta = b * (gen_loss - gen_loss_accum); gen_loss_accum += delta / c
        delta = b * (elbo - elbo_accum); elbo_accum += delta / c
        delta = b * (kl_loss - kl_loss_accum); kl_loss_accum += delta / c
        state = ao.running_means(state, elbo, log_p, kl, b)
    np.testing.assert_allclose(state, [c, elbo_accum, gen_loss_accum, kl_loss_accum], rtol=2e-5)


def test_adam_entry_point_validates_on_host():
    from tvae_b200 import optim
    lib = optim._lib()
    t = (optim.AdamTensor * 1)()
    assert lib.tvae_adam_step(ctypes.cast(t, ctypes.c_void_p), 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, 0, None) < 0   # step is 1-based
    assert b"1-based" in lib.tvae_last_error()
    assert lib.tvae_adam_step(ctypes.cast(t, ctypes.c_void_p), 1, 1e-3, 1.5, 0.999, 1e-8, 0.0, 1, 0, None) < 0
    assert lib.tvae_adam_step(ctypes.cast(t, ctypes.c_void_p), 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 0, None) == 0  # nothing to do
    assert lib.tvae_running_means(None, None, None, 1.0, None, None) < 0


def test_fused_adam_refuses_cpu_parameters():
    from tvae_b200.optim import Adam
    p = torch.nn.Parameter(torch.zeros(4))
    opt = Adam([p], lr=1e-3)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()
    # scheduler compatibility (train_mnist.py:581): ReduceLROnPlateau needs a torch Optimizer with param_groups
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode='max', factor=0.5, patience=0)
    sched.step(1.0); sched.step(0.0)
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-4)
