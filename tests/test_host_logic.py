"""Host-side logic of the drop-in layer on CPU: module/state_dict compatibility, flat parameter views,
normalisation helpers, config surface, loss helpers vs the reference's golden values."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pi-gan-thz_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _models():
    from core.models.discriminator import Discriminator
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    return Generator(250, 4), Discriminator(250, 4), ForwardModel(4, 250, 8)


def test_state_dict_layout_matches_reference_appendix_b():
    G, D, F = _models()
    g = {k: tuple(v.shape) for k, v in G.state_dict().items()}
    assert g["main.0.weight"] == (512, 250) and g["main.3.weight"] == (256, 512) and g["main.6.weight"] == (4, 256)
    assert g["main.1.running_mean"] == (512,) and g["main.4.num_batches_tracked"] == ()
    d = {k: tuple(v.shape) for k, v in D.state_dict().items()}
    assert d == {"main.0.weight": (512, 254), "main.0.bias": (512,), "main.2.weight": (256, 512),
                 "main.2.bias": (256,), "main.4.weight": (1, 256), "main.4.bias": (1,)}
    f = {k: tuple(v.shape) for k, v in F.state_dict().items()}
    assert [k for k in f if k.endswith("weight") and len(f[k]) == 2] == [f"model.{i}.weight" for i in (0, 4, 8, 12, 16, 20)]
    assert f["model.20.weight"] == (258, 256) and f["model.9.weight"] == (1024,)
    assert sum(p.numel() for p in G.parameters()) == 262404
    assert sum(p.numel() for p in D.parameters()) == 262145
    assert sum(p.numel() for p in F.parameters()) == 1385730


def test_reference_state_dicts_load_unchanged():
    from oracle import fixtures
    G, D, F = _models()
    g_sd, d_sd, f_sd = fixtures.make_weights(42)
    G.load_state_dict(g_sd); D.load_state_dict(d_sd); F.load_state_dict(f_sd)   # strict
    assert torch.equal(G.state_dict()["main.3.weight"], g_sd["main.3.weight"])


def test_flat_params_are_views_in_abi_order():
    from pigan_b200 import flat, native
    G, D, F = _models()
    gs, ds, fs = flat.net_state(G, "generator"), flat.net_state(D, "discriminator"), flat.net_state(F, "forward_model")
    assert gs.params.tensor().numel() == native.lib.pigan_generator_param_count(None)
    assert ds.params.tensor().numel() == native.lib.pigan_discriminator_param_count(None)
    assert fs.params.tensor().numel() == native.lib.pigan_forward_model_param_count(None)
    flat_g = gs.params.tensor()
    with torch.no_grad():
        G.main[3].bias.fill_(7.0)
    off = 250 * 512 + 512 + 512 + 512 + 512 * 256
    assert torch.all(flat_g[off:off + 256] == 7.0)          # parameter writes land in the flat buffer
    flat_g[:3] = torch.tensor([1.0, 2.0, 3.0])
    assert G.main[0].weight.view(-1)[:3].tolist() == [1.0, 2.0, 3.0]
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    G.load_state_dict(sd)                                    # load_state_dict copies in place: views survive
    assert gs.params.tensor().data_ptr() == flat_g.data_ptr()
    G.main[1].num_batches_tracked += 2
    assert gs.nbt.tensor().tolist() == [2, 0]
    G.double(); G.float()                                    # _apply re-allocates: views are rebuilt on demand
    assert gs.params.tensor().numel() == 262404 and G.main[3].bias.data_ptr() != 0


def test_modules_refuse_cpu_tensors():
    G, D, F = _models()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            G(torch.zeros(2, 250))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            D(torch.zeros(2, 250), torch.zeros(2, 4))
        F.eval()
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            F(torch.zeros(2, 4))


def test_config_surface():
    import config.config as cfg
    assert (cfg.SPECTRUM_DIM, cfg.Z_DIM, cfg.BATCH_SIZE, cfg.LR_G, cfg.LR_D) == (250, 100, 64, 2e-4, 2e-4)
    assert (cfg.LAMBDA_RECON, cfg.LAMBDA_PHYSICS_SPECTRUM, cfg.LAMBDA_PHYSICS_METRICS, cfg.LAMBDA_MAXWELL, cfg.LAMBDA_LC,
            cfg.LAMBDA_PARAM_RANGE, cfg.LAMBDA_BNN_KL) == (100.0, 10.0, 1.0, 1.0, 1.0, 0.1, 0.0)
    assert cfg.FORWARD_MODEL_OUTPUT_METRICS_DIM == 8 and cfg.SAVE_MODEL_INTERVAL == 50


def test_config_constants_equal_the_reference_module():
    """Every scalar constant of the reference's config/config.py (dumped to tests/golden/config.json in the build
    container) exists in the drop-in module with the same value."""
    import json
    import config.config as cfg
    ref = json.load(open(os.path.join(GOLD, "config.json")))
    assert len(ref) >= 30
    for name, value in ref.items():
        assert hasattr(cfg, name), name
        assert getattr(cfg, name) == value, (name, getattr(cfg, name), value)
    for name in ("CHECKPOINT_DIR", "SAVED_MODELS_DIR", "DATA_DIR", "LOG_DIR", "PLOTS_DIR", "FULL_DATA_PATH", "PROJECT_ROOT"):
        assert isinstance(getattr(cfg, name), str)
    assert callable(cfg.create_directories)


def test_loss_helpers_match_reference_golden():
    from core.utils import loss as L
    from oracle import fixtures
    from oracle import models as O
    g = np.load(os.path.join(GOLD, "forward.npz"))
    _, _, f_sd = fixtures.make_weights(42)
    spec, praw, pnorm, mnorm = fixtures.make_batch(64, seed=7)
    fs, fm = torch.from_numpy(g["f_spec"]), torch.from_numpy(g["f_metrics"])
    np.testing.assert_allclose(L.maxwell_equation_loss(fs, None, pnorm).numpy(), g["loss_maxwell"], rtol=1e-6)
    np.testing.assert_allclose(L.lc_model_approx_loss(fm[:, 0:1], fm[:, 1:2], pnorm).numpy(), g["loss_lc"], rtol=1e-6)
    np.testing.assert_allclose(L.structural_param_range_loss(pnorm * 1.3).numpy(), g["loss_range"], rtol=1e-6)
    assert L.maxwell_equation_loss(torch.zeros(3, 2), None, None).shape == (1,)
    assert L.bnn_kl_loss(torch.nn.Linear(1, 1)).shape == (1,)


def test_dataset_normalisation_and_denormalisation(tmp_path):
    import pandas as pd
    from core.utils import data_loader as dl
    from oracle import fixtures
    spec, praw, pnorm, _ = fixtures.make_batch(32, seed=3)
    freqs = np.linspace(0.5, 3.0, 250)
    cols = {f"Freq_{f:.2f}": spec[:, i].numpy() for i, f in enumerate(freqs)}
    assert len(cols) == 250
    for i, n in enumerate(["r1", "r2", "w", "g"]):
        cols[n] = praw[:, i].numpy()
    rng = np.random.default_rng(0)
    metrics = rng.uniform(0.5, 9.0, size=(32, 8))
    metrics[3, 2] = np.nan
    for i, n in enumerate(["f1", "f2", "Q1", "FoM1", "S1", "Q2", "FoM2", "S2"]):
        cols[n] = metrics[:, i]
    path = tmp_path / "data.csv"
    pd.DataFrame(cols).to_csv(path, index=False)
    ds = dl.MetamaterialDataset(str(path), num_points_per_sample=250)
    assert len(ds) == 32 and len(ds[0]) == 5
    torch.testing.assert_close(ds.normalized_parameters, pnorm, rtol=0, atol=2e-6)
    assert ds.normalized_metrics[3, 2] == 0.5 and float(ds.normalized_metrics.min()) >= 0.0
    back = dl.denormalize_params(ds.normalized_parameters, ds.param_ranges)
    torch.testing.assert_close(back, ds.parameters, rtol=0, atol=1e-6)
    den = dl.denormalize_metrics(ds.normalized_metrics, ds.metric_ranges)
    ok = ~torch.isnan(ds.metrics)
    torch.testing.assert_close(den[ok], ds.metrics[ok], rtol=1e-5, atol=1e-5)
    ns = dl.normalize_spectrum(ds.spectra)
    assert float(ns.min()) == 0.0 and float(ns.max()) == 1.0
    with pytest.raises(FileNotFoundError):
        dl.MetamaterialDataset(str(tmp_path / "missing.csv"))
    assert dl.MetamaterialDataset("", load_data=False).metric_name_to_idx["f2"] == 1


def test_dataset_is_bit_compatible_with_the_reference_class(tmp_path):
    """Drop-in MetamaterialDataset / helpers vs the reference's own class on the same CSV (tests/golden/dataset.npz,
    tools/make_golden.py: dataset): the tensors that feed the step must be identical, NaN pattern included."""
    import pandas as pd
    from core.utils import data_loader as dl
    from oracle import fixtures
    g = np.load(os.path.join(ROOT, "tests", "golden", "dataset.npz"))
    path = tmp_path / "data.csv"
    pd.DataFrame(fixtures.dataset_csv_columns()).to_csv(path, index=False)
    ds = dl.MetamaterialDataset(str(path), num_points_per_sample=250)
    assert len(ds) == int(g["len"])
    np.testing.assert_array_equal(np.asarray(ds.frequencies), g["frequencies"])
    for name in ("spectra", "parameters", "metrics", "normalized_parameters", "normalized_metrics"):
        np.testing.assert_array_equal(np.asarray(getattr(ds, name)), g[name], err_msg=name)
    for k, (lo, hi) in ds.metric_ranges.items():
        np.testing.assert_array_equal(np.array([lo, hi], dtype=np.float64), g[f"range_{k}"], err_msg=k)
    for i, t in enumerate(ds[5]):
        np.testing.assert_array_equal(t.numpy(), g[f"item5_{i}"])
        assert t.dtype == torch.float32
    np.testing.assert_array_equal(
        dl.denormalize_params(torch.as_tensor(ds.normalized_parameters), ds.param_ranges).numpy(), g["denorm_params"])
    np.testing.assert_array_equal(
        dl.denormalize_metrics(torch.as_tensor(ds.normalized_metrics), ds.metric_ranges).numpy(), g["denorm_metrics"])
    np.testing.assert_array_equal(dl.normalize_spectrum(torch.as_tensor(ds.spectra)).numpy(), g["norm_spectrum"])


def test_pretrain_lr_schedule_matches_torch_cosine_annealing():
    """Drop-in pretrain_forward_model applies CosineAnnealingLR(T_max=num_epochs, eta_min=0.01 lr) in closed form
    (pretrain_fwd_model.py:46,135): same per-epoch learning rates as torch's scheduler."""
    import torch.optim as optim
    from torch.optim.lr_scheduler import CosineAnnealingLR
    from core.train.pretrain_fwd_model import cosine_lr
    for T in (1, 3, 10, 500):
        opt = optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
        sch = CosineAnnealingLR(opt, T_max=T, eta_min=1e-5)
        for e in range(min(T, 15)):
            assert abs(opt.param_groups[0]["lr"] - cosine_lr(e, T, 1e-3)) <= 1e-12, (T, e)
            opt.step(); sch.step()


def test_rank_slice_partitions_every_global_batch():
    """device_data.rank_slice: contiguous, disjoint, exhaustive for any (count, world), also ragged last batches."""
    from pigan_b200.device_data import rank_slice
    for count in (0, 1, 5, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                lo, hi = rank_slice(count, r, world)
                assert 0 <= lo <= hi <= count
                covered += list(range(lo, hi))
            assert covered == list(range(count))
