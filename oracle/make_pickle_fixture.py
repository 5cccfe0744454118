"""Generate tests/golden/ref_pickles.pt: whole-module pickles written by the UNMODIFIED reference classes.

The reference checkpoints are `torch.save(model)` pickles of whole modules (src/utils.py:42-46, train_mnist.py:672-684),
resolved by qualified class name (`src.models.<Name>`) when a clustering script loads them (clustering_mnist.py:308);
unpickling does NOT run `__init__`, so the restored objects carry exactly the attributes the reference's constructors
set.  This script builds small modules with the reference constructors (imported from /root/reference), loads the
seeded synthetic weights of the matching golden cases and pickles them; tests/test_reference_pickles.py restores them
with the product's `src.models` on the path.  Run in the build container only:

    python oracle/make_pickle_fixture.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from make_golden import CASES, _import_reference   # noqa: E402
from tvae_b200 import synth                        # noqa: E402

PICKLED = ["g1_mnist", "g8_mnist_attn_unimodal", "g9_mnist_resid", "g12_mnist_attn_unimodal_p4"]


def main():
    ref_models, _, _ = _import_reference()
    out = {}
    for name in PICKLED:
        _, cfg, _ = CASES[name]
        act = nn.Tanh if cfg.activation == "tanh" else nn.LeakyReLU
        with contextlib.redirect_stdout(io.StringIO()):
            gen = ref_models.SpatialGenerator(cfg.z, cfg.hidden, n_out=cfg.n_out, num_layers=cfg.gen_layers, activation=act,
                                              resid=cfg.gen_resid, fourier_expansion=cfg.fourier, sigma=cfg.sigma)
            if cfg.encoder == "attn_unimodal":
                enc = ref_models.InferenceNetwork_AttentionTranslation_UnimodalRotation(
                    cfg.n, cfg.C, cfg.z, kernels_num=cfg.O, activation=act, groupconv=0 if cfg.G == 1 else cfg.G)
            else:
                enc = ref_models.InferenceNetwork_AttentionTranslation_AttentionRotation(
                    cfg.n, cfg.C, cfg.z, kernels_num=cfg.O, kernels_size=cfg.k, padding=cfg.p, activation=act,
                    groupconv=cfg.G, rot_refinement=cfg.rot_refinement, theta_prior=cfg.theta_prior,
                    normal_prior_over_r=cfg.normal_prior_over_r)
        gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state(cfg).items()})
        enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.encoder_state(cfg).items()})
        # the trainers pickle `.eval().cpu()` modules (src/utils.py:42-46)
        bg, be = io.BytesIO(), io.BytesIO()
        torch.save(gen.eval().cpu(), bg)
        torch.save(enc.eval().cpu(), be)
        out[name] = {"generator": bg.getvalue(), "encoder": be.getvalue()}
    path = os.path.join(ROOT, "tests", "golden", "ref_pickles.pt")
    torch.save(out, path)
    print(path, {k: (len(v["generator"]), len(v["encoder"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
