"""Fused `eval_minibatch` for the attention branches (--t-inf attention with --r-inf attention, attention+offsets or
unimodal) - the reference's ELBO call site (train_mnist.py:26-294, train_dsprites.py:27-295, train_galaxy.py:27-295,
train_particles.py:28-343) as one autograd node over the sm_100a kernels.

    elbo, log_p_x_g_z, kl_div = eval_minibatch(x, y, generator_model, encoder_model, t_inf, r_inf, epoch, device,
                                               theta_prior, groupconv, image_dim)
    (-elbo).backward()

Same positional signature, same return triple (0-d tensors on `device` with an autograd graph that populates
`.grad` of every generator / encoder parameter).  `eval_minibatch_particles` takes the particle trainer's
signature (ctf, padding, mask_radius).  Noise: by default drawn on the device from torch's generator with the
distributions the reference uses (Gumbel via -log(Exp(1)), N(0,1)); pass `noise=dict(gumbel, r_z, r_theta)` for
identical-input parity runs.
"""
from __future__ import annotations

import torch

from . import functional as TF


def draw_noise(B, L, z, device, generator=None):
    """Same draws as models.py:387 (gumbel_softmax) and train_mnist.py:206,230."""
    e = torch.empty(B, L, device=device, dtype=torch.float32).exponential_(generator=generator)
    gumbel = -e.log()
    r_z = torch.randn(B, z, device=device, generator=generator)
    r_theta = torch.randn(B, device=device, generator=generator)
    return dict(gumbel=gumbel, r_z=r_z, r_theta=r_theta)


def _encoder_spec(encoder_model, t_inf, r_inf, theta_prior):
    """The inference branches of eval_minibatch that run on the kernels: attention/attention(+offsets)
    (train_mnist.py:187-282) and attention/unimodal with a plain-conv encoder (train_mnist.py:88-183, --groupconv 0)."""
    rotation_attention = hasattr(encoder_model, "rot_refinement")     # ..._AttentionRotation vs ..._UnimodalRotation
    if t_inf == 'attention' and r_inf in ('attention', 'attention+offsets'):
        if not rotation_attention:
            raise ValueError("r_inf attention needs InferenceNetwork_AttentionTranslation_AttentionRotation")
        es = encoder_model.encoder_spec()
        if (r_inf == 'attention+offsets') != es.rot_refinement:
            raise ValueError("r_inf does not match the encoder's rot_refinement")
        return es
    if t_inf == 'attention' and r_inf == 'unimodal':
        if rotation_attention:
            raise ValueError("r_inf unimodal needs InferenceNetwork_AttentionTranslation_UnimodalRotation")
        return encoder_model.encoder_spec(theta_prior)
    raise NotImplementedError("--t-inf unimodal (the MLP encoder of the spatial-VAE baseline) is not on the accelerated hot "
                              "path (SURVEY.md §8)")


def _step(x, y, ctf, generator_model, encoder_model, t_inf, r_inf, device, likelihood, mask_radius, noise, sync=None,
          theta_prior=None, alias=None):
    es = _encoder_spec(encoder_model, t_inf, r_inf, theta_prior)
    spacing = TF.pixel_spacing(x)      # cached on the caller's grid tensor (not on the per-step device copy below)
    y = y.to(device)
    x = x.to(device)
    if not y.is_cuda:
        raise RuntimeError("eval_minibatch: the hot path runs on sm_100a only (no CPU fallback)")
    B = y.shape[0]
    n = y.shape[-1]
    d = n + 2 * es.padding - encoder_model.kernels_size + 1
    if noise is None:
        noise = draw_noise(B, es.attn_G * d * d, es.z, y.device)
    fw, fb = generator_model.fourier_buffers()
    gen_params = generator_model.hot_path_params()
    enc_params = encoder_model.hot_path_params()
    if alias is not None:      # tvae_b200.graph: fresh leaves sharing the parameters' storage (see GraphedStep._eager)
        gen_params = [alias(t) for t in gen_params]
        enc_params = [alias(t) for t in enc_params]
    spec = TF.StepSpec(enc=es, sigma=generator_model._sigma, likelihood=likelihood, mask_radius=int(mask_radius),
                       n_gen_hidden=(len(gen_params) - 5) // 2, gen_resid=bool(generator_model._resid),
                       gen_act=generator_model.act_kind())
    spec.sync = sync
    spec.spacing = spacing
    return TF.FusedStepFn.apply(spec, x, y, ctf, noise["gumbel"], noise["r_z"], noise["r_theta"], fw, fb,
                                *enc_params, *gen_params)


def eval_minibatch(x, y, generator_model, encoder_model, t_inf, r_inf, epoch, device,
                   theta_prior, groupconv, image_dim, noise=None, sync=None, _alias=None):
    """train_mnist / train_dsprites / train_galaxy signature; Bernoulli likelihood (RGB handled by flat order)."""
    return _step(x, y, None, generator_model, encoder_model, t_inf, r_inf, device, "bernoulli", 0, noise, sync, theta_prior, _alias)


def eval_minibatch_particles(x, y, ctf, generator_model, encoder_model, t_inf, r_inf, epoch, device,
                             theta_prior, groupconv, padding, mask_radius, noise=None, sync=None, _alias=None):
    """train_particles signature; Gaussian likelihood with optional CTF and circular mask, or - generator n_out = 2,
    the trainer's --fit-noise (train_particles.py:663-666) - with a learned per-pixel variance.  --fit-noise together
    with a CTF or a mask is rejected: the reference's own shapes do not line up there for B > 1 (y_var becomes
    (B*B, n*n) after the CTF convolution, :304-307; y_var[mask] flattens across the batch, :330-331)."""
    n_out = generator_model.layers[-1].out_features
    if n_out == 2:
        if ctf is not None or int(mask_radius) > 0:
            raise NotImplementedError("--fit-noise with --ctf-file or --mask-radius fails in the reference itself for "
                                      "B > 1 (train_particles.py:304-307, 330-331); not reproduced")
        return _step(x, y, None, generator_model, encoder_model, t_inf, r_inf, device, "gaussian_fit_noise", 0, noise, sync,
                     theta_prior, _alias)
    if n_out != 1:
        raise ValueError(f"eval_minibatch_particles: generator n_out must be 1 or 2 (--fit-noise), got {n_out}")
    if ctf is not None:
        from . import ops
        ops.ctf_filter_size(ctf, y.shape[0], y.shape[-1])     # one odd-sized square filter per image, or raise
        ctf = ctf.to(device)
    return _step(x, y, ctf, generator_model, encoder_model, t_inf, r_inf, device, "gaussian", mask_radius, noise, sync,
                 theta_prior, _alias)


def get_latent(x, y, encoder_model, t_inf, r_inf, device, image_dim, refine=True, return_argmax=False):
    """clustering_*.get_latent, attention branches (clustering_mnist.py:81-120 r_inf unimodal with a plain-conv encoder,
    :122-161 r_inf attention[+offsets]): one encoder pass + one reduction kernel; returns (z_content (B,2z),
    theta_mu (B,1), dx (B,2)) [+ the argmax cell index (B) with return_argmax].

    refine (default): the argmax (r, t) - `attn.view(B,-1).max(1)`, clustering_mnist.py:127 - and the z / theta values
    gathered there are those of an fp32 evaluation: the FP16-operand tensor-core maps only select the near-maximal
    candidate cells, at which the logit chain is re-evaluated with fp32 operands (tvae_refine_argmax).  refine=False
    keeps the fast maps' own argmax (can flip on images whose two best logits are within ~1e-4 of the map's range)."""
    from . import ops
    _encoder_spec(encoder_model, t_inf, r_inf, None if r_inf != 'unimodal' else 1.0)   # validates the branch only
    with torch.no_grad():
        y = y.to(device)
        if not y.is_cuda:
            raise RuntimeError("get_latent: the hot path runs on sm_100a only (no CPU fallback)")
        es = encoder_model.encoder_spec()
        params = encoder_model.hot_path_params()
        heads = TF.encoder_heads_inference(es, y, *params)
        B, NH, G, d, _ = heads.shape
        s = ops.attn_shape(B, G, d, es.z, TF.pixel_spacing(x), es.tables()[1])
        with torch.cuda.device(heads.device):
            zc, th, dx, am = ops.get_latent(s, heads.reshape(B, NH, G, d * d).contiguous())
        if refine:
            r = TF.refine_argmax(es, y, heads, *params)
            zc, th, am = r["z_content"], r["theta_mu"], r["argmax"]
    if return_argmax:
        return zc, th, dx, am
    return zc, th, dx
