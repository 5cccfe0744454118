"""tvae_b200.graph.GraphedStep: the CUDA-graph replay of one training minibatch equals the eager call.

Same models, same minibatch, same generator seed -> the replay draws the noise the eager call draws (torch's graph-safe
Philox offsets), so the ELBO terms agree to fp32 summation order and the gradients to the atomics' order (5e-4 relative
Frobenius; conv_a.bias, whose exact gradient is zero, excluded).  Also: fresh noise per replay, gradients survive
`zero_grad(set_to_none=True)`, the optimiser step after a replay, and the shape / ctf guards."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from test_gpu_step import DEV, build_models, r_inf_of
from tvae_b200 import synth
from tvae_b200.config import CFG1, CFG2, CFG4
from tvae_b200.graph import GraphedStep

pytestmark = pytest.mark.gpu

CASES = [CFG1.with_(name="cfg1_graph", n=24, k=12, p=4), CFG2.with_(name="cfg2_graph", n=32, k=32, p=16),
         CFG4.with_(name="cfg4_graph", n=32, k=16, p=4)]


def make(cfg, B):
    gen, enc = build_models(cfg, seed=3)
    x = torch.from_numpy(synth.image_coords(cfg.n)).to(DEV)
    data = [synth.minibatch(cfg, B, seed=s) for s in (1, 2)]
    particles = cfg.likelihood == "gaussian"
    y = [torch.from_numpy(d["y"]).to(DEV) for d in data]
    ctf = [torch.from_numpy(d["ctf"]).to(DEV) if particles and d["ctf"] is not None else None for d in data]
    kw = dict(particles=True, padding=cfg.p, mask_radius=cfg.mask_radius, ctf_shape=None if ctf[0] is None else ctf[0].shape) if particles else {}
    gs = GraphedStep(x, y[0].shape, gen, enc, "attention", r_inf_of(cfg), DEV, cfg.theta_prior, cfg.G, cfg.n, **kw)
    return gen, enc, x, y, ctf, gs


def eager(cfg, gen, enc, x, y, ctf):
    from tvae_b200 import elbo as E
    for p in list(gen.parameters()) + list(enc.parameters()):
        p.grad = None
    if cfg.likelihood == "gaussian":
        out = E.eval_minibatch_particles(x, y, ctf, gen, enc, "attention", r_inf_of(cfg), 0, DEV, cfg.theta_prior, cfg.G, cfg.p,
                                         cfg.mask_radius)
    else:
        out = E.eval_minibatch(x, y, gen, enc, "attention", r_inf_of(cfg), 0, DEV, cfg.theta_prior, cfg.G, cfg.n)
    (-out[0]).backward()
    names = [f"gen.{n}" for n, _ in gen.named_parameters()] + [f"enc.{n}" for n, _ in enc.named_parameters()]
    grads = {n: p.grad.detach().clone() for n, p in zip(names, list(gen.parameters()) + list(enc.parameters()))}
    return [float(t.detach()) for t in out], grads


@pytest.mark.parametrize("cfg", CASES, ids=lambda c: c.name)
def test_replay_equals_eager(cfg):
    B = 6
    gen, enc, x, y, ctf, gs = make(cfg, B)
    for i in (1, 0):                       # the second minibatch first: the staging buffers are really overwritten
        torch.manual_seed(77 + i)
        terms_e, grads_e = eager(cfg, gen, enc, x, y[i], ctf[i])
        torch.manual_seed(77 + i)
        out = gs(y[i], ctf[i])
        terms_g = [float(t) for t in out]
        np.testing.assert_allclose(terms_g, terms_e, rtol=2e-5, atol=1e-4)
        names = list(grads_e)
        for n, p in zip(names, gs.params):
            if n == "enc.conv_a.bias":
                continue
            assert p.grad is not None
            assert rel_err(p.grad, grads_e[n]) < 5e-4, n


def test_replays_draw_fresh_noise_and_keep_gradients():
    cfg, B = CASES[0], 6
    gen, enc, x, y, ctf, gs = make(cfg, B)
    e1 = float(gs(y[0])[0])
    e2 = float(gs(y[0])[0])
    assert e1 != e2                        # same images, new Gumbel / normal draws
    opt = torch.optim.SGD(gs.params, lr=0.0)
    opt.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in gs.params)
    gs(y[1])
    assert all(p.grad is g for p, g in zip(gs.params, gs.grads))
    # a real optimiser step between replays (parameters change in place; the graph reads them where they are)
    from tvae_b200.optim import Adam
    adam = Adam(gs.params, lr=1e-3)
    torch.manual_seed(5)
    before = float(gs(y[0])[0])
    adam.step()
    torch.manual_seed(5)
    after = float(gs(y[0])[0])
    assert after != before and np.isfinite(after)


def test_guards():
    cfg, B = CASES[0], 4
    gen, enc, x, y, ctf, gs = make(cfg, B)
    with pytest.raises(ValueError):
        gs(y[0][:2])
    with pytest.raises(ValueError):
        gs(y[0], torch.zeros(B, 1, cfg.n - 1, cfg.n - 1, device=DEV))
    with pytest.raises(RuntimeError):
        GraphedStep(x, y[0].shape, gen, enc, "attention", r_inf_of(cfg), "cpu", cfg.theta_prior, cfg.G, cfg.n)


def test_train_epoch_with_graph_matches_eager():
    """train_epoch(graph=True): three full minibatches replayed, the last shorter one eager; same seeds -> the parameters end
    where the eager epoch puts them (Adam moves a weight by ~lr per step: same bound as tests/test_gpu_optim.py)."""
    from tvae_b200 import train
    from tvae_b200.optim import Adam
    cfg, B = CASES[0], 6
    x = torch.from_numpy(synth.image_coords(cfg.n)).to(DEV)
    batches = [(torch.from_numpy(synth.minibatch(cfg, B if i < 3 else 4, seed=i)["y"]).to(DEV),) for i in range(4)]
    finals, outs = [], []
    for graph in (False, True):
        gen, enc = build_models(cfg, seed=3)
        params = list(gen.parameters()) + list(enc.parameters())
        opt = Adam(params, lr=2e-4)
        cache = {} if graph else False
        if graph:      # capture (with its warm-up passes, which draw noise) before seeding, so both epochs see the same stream
            train.train_epoch(batches[:1], x, gen, enc, Adam(params, lr=0.0), "attention", r_inf_of(cfg), 0, 1, B, DEV, params,
                              cfg.theta_prior, cfg.G, cfg.n, graph=cache)
            assert len(cache) == 1
        torch.manual_seed(99)
        out = train.train_epoch(batches, x, gen, enc, opt, "attention", r_inf_of(cfg), 0, 1, 3 * B + 4, DEV, params,
                                cfg.theta_prior, cfg.G, cfg.n, graph=cache)
        torch.cuda.synchronize()
        assert all(np.isfinite(out))
        assert all(int(opt.state[p]["step"]) == 4 for p in params)
        finals.append([p.detach().clone() for p in params])
        outs.append(out)
    np.testing.assert_allclose(outs[1], outs[0], rtol=1e-3)
    lr, steps = 2e-4, 4
    d = torch.cat([(a - b).abs().flatten() for a, b in zip(*finals)])
    assert float(d.max()) <= 2 * lr * steps * 1.01 and float(d.mean()) <= 0.02 * lr
