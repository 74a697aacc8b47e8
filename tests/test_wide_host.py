"""Host-side checks of the widened dims (BASELINE config 5) that need no GPU: which PiganDims the library accepts,
the flat parameter count against the oracle's state_dict, and the drop-in module's constructor."""
import ctypes as C
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pi-gan-thz_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)


def test_accepted_dims():
    from pigan_b200 import native
    lib = native.lib
    ok = native.make_dims(spectrum_dim=2048, f_hidden=(2048,) * 5, g_hidden=(2048, 2048), d_hidden=(2048, 2048))
    assert lib.pigan_engine_workspace_bytes(C.byref(ok), 1024) > 0
    mixed = native.make_dims(spectrum_dim=500, f_hidden=(512, 1024, 2048, 1024, 256))
    assert lib.pigan_engine_workspace_bytes(C.byref(mixed), 1024) > 0
    assert lib.pigan_engine_workspace_bytes(C.byref(native.default_dims()), 1024) > 0
    for bad in (native.make_dims(f_hidden=(256, 512, 768, 512, 256)),      # width not 256/512/1024/2048
                native.make_dims(spectrum_dim=251),                         # odd spectrum length
                native.make_dims(spectrum_dim=4096, f_hidden=(2048,) * 5)):  # S + Mt beyond 2560
        assert lib.pigan_engine_workspace_bytes(C.byref(bad), 1024) == 0
    # a surrogate-only engine (spectrum length not a multiple of 64: no PI-GAN step) carries no generator /
    # discriminator activations; the config-5 engine carries the whole step's
    assert lib.pigan_engine_workspace_bytes(C.byref(mixed), 65536) < 3 << 30
    assert 4 << 30 < lib.pigan_engine_workspace_bytes(C.byref(ok), 65536) < 16 << 30


def test_param_count_matches_the_oracle_layout():
    from oracle import models as O
    from pigan_b200 import native
    hidden = (2048,) * 5
    dims = native.make_dims(spectrum_dim=2048, f_hidden=hidden)
    sd = O.init_forward_model(4, 2048, 8, hidden, gen=torch.Generator().manual_seed(0))
    n = sum(v.numel() for k, v in sd.items())
    assert native.lib.pigan_forward_model_param_count(C.byref(dims)) == n


def test_module_takes_widths():
    from core.models.forward_model import ForwardModel
    F = ForwardModel(4, 2048, 8, hidden=(2048,) * 5)
    assert F.model[0].out_features == 2048 and F.model[20].out_features == 2056
    assert list(F.state_dict()) == list(ForwardModel(4, 250, 8).state_dict())     # same keys as the reference stack
    assert ForwardModel(4, 250, 8).engine_dims() is None
    d = F.engine_dims()
    assert d.spectrum_dim == 2048 and tuple(d.f_hidden) == (2048,) * 5
    with pytest.raises(ValueError):
        ForwardModel(4, 250, 8, hidden=(256, 512))


def test_trainer_reads_the_dims_off_the_modules():
    from core.models.discriminator import Discriminator
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    from pigan_b200 import native
    from pigan_b200.trainer import NativeTrainer, dp_phase_plan
    assert NativeTrainer._dims_of(Generator(250, 4), Discriminator(250, 4), ForwardModel(4, 250, 8)) is None
    G, D = Generator(2048, 4, hidden=(2048, 2048)), Discriminator(2048, 4, hidden=(2048, 2048))
    F = ForwardModel(4, 2048, 8, hidden=(2048,) * 5)
    d = NativeTrainer._dims_of(G, D, F)
    assert (d.spectrum_dim, d.metrics_dim, tuple(d.g_hidden), tuple(d.d_hidden)) == (2048, 8, (2048, 2048), (2048, 2048))
    assert list(G.state_dict()) == list(Generator(250, 4).state_dict())
    assert list(D.state_dict()) == list(Discriminator(250, 4).state_dict())
    # flat layouts of the widened generator / discriminator = the modules' parameter counts
    assert native.lib.pigan_generator_param_count(C.byref(d)) == sum(p.numel() for p in G.parameters())
    assert native.lib.pigan_discriminator_param_count(C.byref(d)) == sum(p.numel() for p in D.parameters())
    with pytest.raises(ValueError):
        NativeTrainer._dims_of(G, Discriminator(250, 4), F)
    # the data-parallel plan exchanges BatchNorm slices of the widened widths
    plan = dict(dp_phase_plan(2048, 2048))
    assert plan[0] == [("bn_sums", slice(0, 4096))] and plan[1] == [("bn_sums", slice(4096, 8192))]


def test_bench_flop_formulas_reproduce_the_reference_width_constants():
    """bench.py's per-sample FLOP formulas for the widened blocks are the ones behind the reference-width constants
    (SURVEY 8(d)): evaluated at the reference dims they must give exactly those."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    ref_f = (256, 512, 1024, 512, 256)
    assert bench.wide_gan_flop_per_sample(250, 4, 8, (512, 256), (512, 256), ref_f) == bench.FLOP_PER_TRAIN_SAMPLE
    assert bench.wide_flop_per_sample(250, 8, ref_f) == bench.FLOP_PER_PRETRAIN_SAMPLE
    assert bench.wide_gan_flop_per_sample() == 193_167_360       # BASELINE config 5: "193 MFLOP per sample"
    assert bench.wide_flop_per_sample() == 125_960_192


def test_closed_form_layernorm_statistics_of_the_k4_layer():
    """csrc/elementwise.cu: f_l1_consts_kernel / f_l1_wide_kernel.  For h = W p + b with K = 4 the row statistics over
    the N columns are closed-form in p: mean = mean_c(W) . p + mean(b), var = p^T A p + 2 p^T B + C with the centred
    second moments A, B, C (fp64 constants, fp32 evaluation).  Restated in numpy and held to the direct two-pass
    statistics, also when a large common offset would make E[h^2] - mean^2 cancel."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(0))
    for N, offset in ((256, 0.0), (2048, 0.0), (2048, 300.0)):
        W = rng.uniform(-0.5, 0.5, size=(N, 4)).astype(np.float32)
        b = (rng.uniform(-0.5, 0.5, size=N) + offset).astype(np.float32)
        p = rng.uniform(-1, 1, size=(64, 4)).astype(np.float32)
        Wd, bd = W.astype(np.float64), b.astype(np.float64)
        U, v = Wd - Wd.mean(0), bd - bd.mean()
        A, Bv, Cc = (U.T @ U) / N, (U * v[:, None]).mean(0), (v * v).mean()
        A32, B32, C32 = A.astype(np.float32), Bv.astype(np.float32), np.float32(Cc)
        var_cf = np.einsum("ri,ij,rj->r", p, A32, p) + 2 * (p @ B32) + C32            # fp32, as the row kernel does
        h = p.astype(np.float64) @ Wd.T + bd
        var_ref = h.var(axis=1)
        assert np.all(var_cf > 0)
        np.testing.assert_allclose(var_cf, var_ref, rtol=2e-5)
        xhat_cf = (p @ U.astype(np.float32).T + v.astype(np.float32)) / np.sqrt(var_cf + 1e-5)[:, None]
        xhat_ref = (h - h.mean(1, keepdims=True)) / np.sqrt(var_ref + 1e-5)[:, None]
        assert np.abs(xhat_cf - xhat_ref).max() < 2e-5
