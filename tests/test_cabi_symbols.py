"""The C-ABI library loads on a CPU-only box and exports every symbol include/tvae_b200.h declares
(no compute calls without a GPU).  Also checks host-side argument validation that needs no device."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tvae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tvae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from tvae_b200 import _lib
    lib = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tvae_b200.h but not exported"
    lib.tvae_version.restype = ctypes.c_int
    assert lib.tvae_version() >= 100


def test_host_side_validation_without_gpu():
    from tvae_b200 import ops
    lib = ops.L()
    assert lib.tvae_bank_pitch(1, 28) == 800 and lib.tvae_bank_pitch(1, 64) == 4128
    bad = ops.EncShape(2, 1, 20, 9, 3, 8, 33, 2, lib.tvae_bank_pitch(1, 9))     # O not a multiple of 32
    rc = lib.tvae_filter_bank_fwd(ctypes.byref(bad), None, None, None)
    assert rc < 0 and b"multiple of 32" in lib.tvae_last_error()


def test_product_path_fails_loudly_on_cpu_tensors():
    import contextlib, io
    import pytest
    import torch
    import src.models as models
    with contextlib.redirect_stdout(io.StringIO()):
        conv = models.GroupConv(1, 32, 9, padding=3, output_rot_dim=8)
        gen = models.SpatialGenerator(2, 64, num_layers=2, fourier_expansion=True, sigma=0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv(torch.zeros(1, 1, 20, 20), "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gen(torch.zeros(1, 4, 2), torch.zeros(1, 2))


def test_state_dict_contract():
    """Parameter names/shapes of SURVEY.md §8b (reference checkpoints must load)."""
    import contextlib, io
    import src.models as models
    with contextlib.redirect_stdout(io.StringIO()):
        gen = models.SpatialGenerator(2, 512, num_layers=2, fourier_expansion=True, sigma=2 / 49)
        enc = models.InferenceNetwork_AttentionTranslation_AttentionRotation(
            50, 1, 2, kernels_num=128, kernels_size=28, padding=8, groupconv=8, rot_refinement=True,
            normal_prior_over_r=False)
    sd = {k: tuple(v.shape) for k, v in gen.state_dict().items()}
    assert sd == {"embed_latent.weight": (1024, 2), "embed_latent.bias": (1024,), "coord_linear.weight": (512, 1024),
                  "coord_linear.bias": (512,), "latent_linear.weight": (512, 2), "layers.1.weight": (512, 512),
                  "layers.1.bias": (512,), "layers.3.weight": (1, 512), "layers.3.bias": (1,)}
    se = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    assert se["conv1.weight"] == (128, 1, 1, 28, 28) and se["conv2.weight"] == (128, 128, 1, 1, 1)
    assert se["conv_a.weight"] == (1, 128, 1, 1, 1) and se["conv_r.weight"] == (2, 128, 1, 1, 1)
    assert se["conv_z.weight"] == (4, 128, 1, 1, 1)
    n_params = sum(p.numel() for p in gen.parameters()) + sum(p.numel() for p in enc.parameters())
    assert n_params == 906888
