"""Training / evaluation epochs with the reference's call convention (train_mnist.py:296-392, train_particles.py:
345-450) on the fused hot path: eval_minibatch -> backward -> one-launch Adam, running means on the device.

    elbo, gen_loss, kl_loss = train_epoch(iterator, x_coord, generator_model, encoder_model, optim, t_inf, r_inf,
                                          epoch, num_epochs, N, device, params, theta_prior, groupconv, image_dim)

`optim` is `tvae_b200.optim.Adam` (fused step) or any torch optimiser.  Batches are `(y,)` or `(y, ctf)` tuples like
the reference's DataLoaders yield; the particle trainer's `padding` / `mask_radius` are keyword arguments.
`graph=True` replays forward + backward of every full-size minibatch as one CUDA graph (tvae_b200.graph.GraphedStep, built
on the first minibatch of a shape and kept in the dict passed as `graph`, or in a fresh one per epoch for `graph=True`);
a last, shorter minibatch and a `sync` with a fused optimiser take the eager path.
"""
from __future__ import annotations

import sys

import torch

from . import elbo as E
from .optim import Adam, RunningMeans


def _minibatch(x, batch, generator_model, encoder_model, t_inf, r_inf, epoch, device, theta_prior, groupconv, image_dim,
               particles, padding, mask_radius, sync=None):
    if particles:
        y, ctf = (batch[0], batch[1] if len(batch) > 1 else None)
        return y.size(0), E.eval_minibatch_particles(x, y, ctf, generator_model, encoder_model, t_inf, r_inf, epoch, device,
                                                      theta_prior, groupconv, padding, mask_radius, sync=sync)
    y = batch[0]
    return y.size(0), E.eval_minibatch(x, y, generator_model, encoder_model, t_inf, r_inf, epoch, device, theta_prior,
                                       groupconv, image_dim, sync=sync)


def train_epoch(iterator, x_coord, generator_model, encoder_model, optim, t_inf, r_inf, epoch, num_epochs, N, device, params,
                theta_prior, groupconv, image_dim, particles=False, padding=0, mask_radius=0, sync=None, progress=False,
                graph=False):
    generator_model.train()
    encoder_model.train()
    stats = RunningMeans(device)
    fused = isinstance(optim, Adam)
    graphs = graph if isinstance(graph, dict) else ({} if graph else None)
    if graphs is not None and sync is not None and getattr(sync, "optimizer", None) is not None:
        graphs = None                     # the fused optimiser epilogue carries a host-side step count: eager path
    c = 0
    for batch in iterator:
        step = None
        if graphs is not None:
            y = batch[0]
            ctf = batch[1] if particles and len(batch) > 1 else None
            key = (tuple(y.shape), None if ctf is None else tuple(ctf.shape))
            step = graphs.get(key)
            if step is None and (not graphs or y.shape[0] >= max(k[0][0] for k in graphs)):
                from .graph import GraphedStep
                step = graphs[key] = GraphedStep(x_coord, y.shape, generator_model, encoder_model, t_inf, r_inf, device, theta_prior,
                                                 groupconv, image_dim, ctf_shape=None if ctf is None else tuple(ctf.shape),
                                                 particles=particles, padding=padding,
                                                 mask_radius=mask_radius, sync=sync)
        if step is not None:
            b = batch[0].size(0)
            elbo, log_p_x_g_z, kl_div = step(batch[0], batch[1] if particles and len(batch) > 1 else None)
        else:
            if graphs:
                optim.zero_grad(set_to_none=True)      # detach a graph's static gradients: the eager backward must not add to them
            b, (elbo, log_p_x_g_z, kl_div) = _minibatch(x_coord, batch, generator_model, encoder_model, t_inf, r_inf, epoch, device,
                                                         theta_prior, groupconv, image_dim, particles, padding, mask_radius, sync)
            loss = -elbo
            loss.backward()
        if step is not None:
            optim.step()       # the gradients are the graph's static tensors: they stay attached, the next replay overwrites them
        elif sync is not None and getattr(sync, "optimizer", None) is optim:
            # the update already ran inside the backward pass, bucket by bucket, behind each gradient all-reduce
            optim.zero_grad(set_to_none=True)
        elif fused:
            optim.step()
            optim.zero_grad(set_to_none=True)
        else:
            optim.step()
            optim.zero_grad()
        stats.update(elbo.detach(), log_p_x_g_z.detach(), kl_div.detach(), b)
        c += b
        if progress:   # the reference prints running values every step (three host syncs); opt-in here
            e, g, k = stats.read()
            print('# [{}/{}] training {:.1%}, ELBO={:.5f}, Error={:.5f}, KL={:.5f}'.format(epoch + 1, num_epochs, c / N, e, g, k),
                  end='\r', file=sys.stderr)
    return stats.read()


def eval_model(iterator, x_coord, generator_model, encoder_model, t_inf, r_inf, epoch, device, theta_prior, groupconv,
               image_dim, particles=False, padding=0, mask_radius=0):
    generator_model.eval()
    encoder_model.eval()
    stats = RunningMeans(device)
    with torch.no_grad():
        for batch in iterator:
            b, (elbo, log_p_x_g_z, kl_div) = _minibatch(x_coord, batch, generator_model, encoder_model, t_inf, r_inf, epoch,
                                                         device, theta_prior, groupconv, image_dim, particles, padding,
                                                         mask_radius)
            stats.update(elbo, log_p_x_g_z, kl_div, b)
    return stats.read()
