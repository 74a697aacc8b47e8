"""On-device data pipeline (pigan_gather_rows, pigan_generate_spectra).
Gather: bit-exact against torch indexing.  Generator: fp32 against the reference's own function
(tests/golden/datagen.npz, produced by generate_single_terahertz_spectrum_and_params) and the float64 oracle —
absolute 2e-5 on values of magnitude <= 14 (fp32 exp / tanh)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pi-gan-thz_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"
TOL_ABS = 2e-5


@pytest.mark.parametrize("shape,dtype", [((1000, 256), torch.float16), ((777, 250), torch.float32),
                                          ((513, 8), torch.float32), ((64, 4), torch.float32),
                                          ((100, 3), torch.float16), ((50, 7), torch.uint8), ((33,), torch.float32)])
def test_gather_rows_is_bit_exact(shape, dtype):
    from pigan_b200 import device_data as DD
    g = torch.Generator(device=DEV).manual_seed(1)
    src = (torch.rand(shape, device=DEV, generator=g) * 200).to(dtype)
    for count in (0, 1, 31, shape[0], 3 * shape[0] + 5):
        idx = torch.randint(0, shape[0], (count,), device=DEV, generator=g)
        out = DD.gather_rows(src, idx)
        assert out.shape[0] == count and torch.equal(out, src[idx])


def test_gather_rows_flags_bad_indices_instead_of_reading_out_of_bounds():
    from pigan_b200 import native
    src = torch.arange(40, device=DEV, dtype=torch.float32).view(10, 4)
    idx = torch.tensor([0, 9, 10, -1, 3], device=DEV)
    out = torch.full((5, 4), 7.0, device=DEV)
    bad = torch.zeros(1, device=DEV, dtype=torch.int32)
    native.check(native.lib.pigan_gather_rows(src.data_ptr(), 10, 16, idx.data_ptr(), 5, out.data_ptr(), bad.data_ptr(),
                                              native.current_stream()))
    assert int(bad) == 1
    assert torch.equal(out[[0, 1, 4]], src[[0, 9, 3]]) and float(out[2:4].abs().sum()) == 0.0


def test_generator_matches_reference_function():
    from oracle import datagen as DG
    from pigan_b200 import device_data as DD
    g = np.load(os.path.join(GOLD, "datagen.npz"))
    params = torch.from_numpy(g["params"]).float().to(DEV)
    freq64 = np.linspace(0.5, 3.0, 250)
    freq = torch.from_numpy(freq64).float().to(DEV)
    n = params.shape[0]
    clean, p_back = DD.generate_spectra(n, DEV, noise_level=0.0, params_denorm=params, frequency=freq)
    assert torch.equal(p_back, params)
    assert np.max(np.abs(clean.cpu().numpy() - g["clean"])) < TOL_ABS
    nooff, _ = DD.generate_spectra(n, DEV, noise_level=0.0, params_denorm=params, frequency=freq, apply_offset=False)
    assert np.max(np.abs(nooff.cpu().numpy() - g["clean_nooffset"])) < TOL_ABS
    # with noise: replay the kernel's own z through the oracle (numpy's RNG stream is not reproduced)
    dump = torch.empty(n, 250, device=DEV)
    noisy, _ = DD.generate_spectra(n, DEV, seed=5, noise_level=0.1, params_denorm=params, frequency=freq,
                                   noise_dump=dump)
    ref = DG.generate_spectra(freq64, params.cpu().numpy().astype(np.float64), noise=dump.cpu().numpy(),
                              noise_level=0.1)
    assert np.max(np.abs(noisy.cpu().numpy() - ref)) < TOL_ABS
    assert float(noisy.max()) <= 0.0


def test_generator_noise_and_parameters_are_counter_based():
    from pigan_b200 import device_data as DD
    n = 4096
    dump = torch.empty(n, 250, device=DEV)
    spec, p = DD.generate_spectra(n, DEV, seed=9, noise_level=0.1, noise_dump=dump)
    # statistics of the draws
    assert abs(float(dump.mean())) < 5e-3 and abs(float(dump.std()) - 1.0) < 5e-3
    assert float(p.min()) >= 2.2 and float(p.max()) <= 2.8 and abs(float(p.mean()) - 2.5) < 5e-3
    # any batching / sharding of the same global rows gives the same data
    a, pa = DD.generate_spectra(1000, DEV, seed=9, first_index=0, noise_level=0.1)
    b, pb = DD.generate_spectra(n - 1000, DEV, seed=9, first_index=1000, noise_level=0.1)
    assert torch.equal(torch.cat([a, b]), spec) and torch.equal(torch.cat([pa, pb]), p)
    c, _ = DD.generate_spectra(n, DEV, seed=10, noise_level=0.1)
    assert not torch.equal(c, spec)
    # peak indices of generated rows are what the physics kernel reports for them (bit-exact argmin)
    from pigan_b200 import native, synthetic
    idx = torch.empty(n, device=DEV, dtype=torch.int32)
    out = torch.empty(n, 4, device=DEV)
    freq = synthetic.frequencies(250, device=DEV)
    native.check(native.lib.pigan_physics_metrics(spec.data_ptr(), n, 250, freq.data_ptr(), None, 0.0, idx.data_ptr(),
                                                  out.data_ptr(), native.current_stream()))
    assert torch.equal(idx.long(), spec.argmin(dim=1))


def test_device_loader_epochs_cover_the_dataset_and_shard_across_ranks():
    from pigan_b200 import device_data as DD
    ds = DD.DeviceDataset.synthetic(1000, DEV, seed=3)
    assert len(ds) == 1000
    ld = DD.DeviceLoader(ds, batch_size=96, shuffle=True, seed=1)
    assert len(ld) == 11 and ld.batch_size == 96
    rows = []
    for spec, pden, pnorm, mden, mnorm in ld:
        assert spec.is_cuda and spec.shape[1] == 250 and pden.shape[1] == 4 and mnorm.shape[1] == 8
        assert torch.allclose(pnorm, (pden - 2.2) / 0.6 * 2 - 1)
        rows.append(spec)
    got = torch.cat(rows)
    assert got.shape[0] == 1000
    # every dataset row exactly once (rows are distinct: noisy spectra)
    key = lambda t: sorted(map(float, t.double().sum(dim=1).cpu()))
    assert key(got) == key(ds.spectra)
    # next epoch: another order; same epoch + seed on another loader object: same order
    first2 = next(iter(ld))[0]
    assert not torch.equal(first2, rows[0])
    again = next(iter(DD.DeviceLoader(ds, batch_size=96, shuffle=True, seed=1)))[0]
    assert torch.equal(again, rows[0])
    # two ranks split every global batch; together they see what one rank with the global batch sees
    one = list(DD.DeviceLoader(ds, 64, seed=4).batch_indices(0))
    r0 = list(DD.DeviceLoader(ds, 32, seed=4, rank=0, world=2).batch_indices(0))
    r1 = list(DD.DeviceLoader(ds, 32, seed=4, rank=1, world=2).batch_indices(0))
    assert len(one) == len(r0) == len(r1)
    for a, b, c in zip(one, r0, r1):
        assert torch.equal(a, torch.cat([b, c]))
    assert len(DD.DeviceLoader(ds, 96, drop_last=True)) == 10


def test_reference_trainers_run_from_the_device_loader(tmp_path):
    """train_pigan and pretrain_forward_model (drop-in modules) take the DeviceLoader where the reference takes a
    DataLoader; nothing crosses PCIe per batch."""
    import config.config as cfg
    from core.models.discriminator import Discriminator
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    from core.train.pretrain_fwd_model import pretrain_forward_model
    from core.train.train_pigan import train_pigan
    from pigan_b200 import device_data as DD
    cfg.SAVED_MODELS_DIR = str(tmp_path / "saved")
    cfg.CHECKPOINT_DIR = str(tmp_path / "ckpt")
    ds = DD.DeviceDataset.synthetic(512, DEV, seed=8)
    F = ForwardModel(4, 250, 8)
    hist = pretrain_forward_model(F, DD.DeviceLoader(ds, 128, seed=2), torch.device(DEV), num_epochs=3, lr=1e-3)
    assert len(hist) == 3 and hist[-1] < hist[0]

    class Meta:   # what train_pigan reads from the dataset object (train_pigan.py:132,162,165-166)
        param_ranges = {k: (2.2, 2.8) for k in ("r1", "r2", "w", "g")}
        frequencies = np.linspace(0.5, 3.0, 250)
        metric_name_to_idx = {"f1": 0, "f2": 1}
    F.eval()
    lh = train_pigan(DD.DeviceLoader(ds, 128, seed=3), torch.device(DEV), Generator(250, 4), Discriminator(250, 4), F,
                     ds, num_epochs=2, log_interval=10)     # the synthetic dataset carries what train_pigan reads
    assert Meta.metric_name_to_idx == {k: ds.metric_name_to_idx[k] for k in ("f1", "f2")}
    assert len(lh["g_losses"]) == 2 and all(np.isfinite(v) for v in lh["g_losses"] + lh["d_losses"])


def test_peak_shift_is_the_difference_of_argmin_frequencies():
    """physics.peak_shift (SURVEY 8(f) N2's definition): indices bit-exact against numpy argmin, shift 1e-6."""
    from pigan_b200 import device_data as DD
    from pigan_b200 import physics
    a, _ = DD.generate_spectra(2048, DEV, seed=21, noise_level=0.1)
    b, _ = DD.generate_spectra(2048, DEV, seed=22, noise_level=0.1)
    r = physics.peak_shift(a, b)
    freq = np.linspace(0.5, 3.0, 250)
    ia, ib = a.cpu().numpy().argmin(axis=1), b.cpu().numpy().argmin(axis=1)
    assert np.array_equal(r["peak_idx_reconstructed"].cpu().numpy(), ia)
    assert np.array_equal(r["peak_idx_target"].cpu().numpy(), ib)
    np.testing.assert_allclose(r["peak_shift"].cpu().numpy(), freq[ia] - freq[ib], atol=1e-6)
    assert float(physics.peak_shift(a, a)["peak_shift"].abs().max()) == 0.0
    m = physics.peak_metrics(a)
    ok = ~torch.isnan(m["Q"])
    assert ok.float().mean() > 0.9 and torch.allclose(m["S"][ok], m["f_res"][ok] * m["Q"][ok], rtol=1e-5)


def test_checkpoint_files_have_the_reference_layout(tmp_path):
    """Files and their structure as written by the reference's train_pigan (train_pigan.py:284-309; golden
    tests/golden/checkpoint_structure.json from a reference run with SAVE_MODEL_INTERVAL = 1): same file names,
    checkpoint keys, state_dict keys / shapes / dtypes, Adam state entries and param-group hyper-parameters."""
    import json
    import config.config as cfg
    from core.models.discriminator import Discriminator
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    from core.train.train_pigan import train_pigan
    from oracle import fixtures
    gold = json.load(open(os.path.join(GOLD, "checkpoint_structure.json")))
    cfg.SAVED_MODELS_DIR, cfg.CHECKPOINT_DIR = str(tmp_path / "saved"), str(tmp_path / "ckpt")
    old = cfg.SAVE_MODEL_INTERVAL
    cfg.SAVE_MODEL_INTERVAL = 1

    class Meta:
        param_ranges = {k: (2.2, 2.8) for k in ("r1", "r2", "w", "g")}
        frequencies = np.linspace(0.5, 3.0, 250)
        metric_name_to_idx = {"f1": 0, "f2": 1}
    try:
        spec, praw, pnorm, mnorm = fixtures.make_batch(64, seed=100)
        F = ForwardModel(4, 250, 8)
        F.eval()
        train_pigan([(spec, praw, pnorm, torch.zeros(64, 8), mnorm)], torch.device(DEV), Generator(250, 4),
                    Discriminator(250, 4), F, Meta(), num_epochs=1, log_interval=10)
    finally:
        cfg.SAVE_MODEL_INTERVAL = old
    assert sorted(os.listdir(cfg.CHECKPOINT_DIR)) == gold["checkpoint_files"]
    assert sorted(os.listdir(cfg.SAVED_MODELS_DIR)) == gold["saved_files"]

    def structure(obj):
        if isinstance(obj, dict):
            return {str(k): structure(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return [structure(v) for v in obj]
        if isinstance(obj, torch.Tensor):
            return {"tensor": list(obj.shape), "dtype": str(obj.dtype)}
        if isinstance(obj, (bool, int, float)) or obj is None:
            return {"scalar": type(obj).__name__}
        return {"other": type(obj).__name__}

    ck = torch.load(os.path.join(cfg.CHECKPOINT_DIR, "pigan_epoch_1.pth"), map_location="cpu", weights_only=False)
    got = structure(ck)
    assert set(got) == set(gold["checkpoint"])
    for key in ("generator_state_dict", "discriminator_state_dict", "forward_model_state_dict"):
        assert got[key] == gold["checkpoint"][key], key
    for key in ("optimizer_g_state_dict", "optimizer_d_state_dict"):
        assert got[key]["state"] == gold["checkpoint"][key]["state"], key          # per-parameter step / exp_avg / exp_avg_sq
        assert set(got[key]["param_groups"][0]) == set(gold["checkpoint"][key]["param_groups"][0]), key
    pg = ck["optimizer_g_state_dict"]["param_groups"][0]
    for k, v in gold["param_group_g"].items():
        if k == "lr":
            assert abs(pg[k] - v) <= 1e-12, (k, pg[k], v)       # CosineAnnealingLR after one epoch of one
        else:
            assert (list(pg[k]) if isinstance(pg[k], (list, tuple)) else pg[k]) == v, (k, pg[k], v)
    hist = torch.load(os.path.join(cfg.SAVED_MODELS_DIR, "pigan_loss_history.pt"), weights_only=False)
    assert structure(hist) == gold["loss_history"]
    # and the reference's modules can load what was written
    assert set(torch.load(os.path.join(cfg.SAVED_MODELS_DIR, "generator_final.pth"), map_location="cpu").keys()) == \
        set(gold["checkpoint"]["generator_state_dict"])


def test_differentiable_physics_metrics_match_the_autograd_oracle():
    """pigan_physics_metrics_backward against oracle/physics.py: peak_parameters_vjp (torch float64
    autograd on the reference's arithmetic with its branch decisions): fp32 outputs of fp64 math, 1e-5 relative."""
    from oracle import fixtures
    from oracle import physics as P
    from pigan_b200 import physics
    spec, *_ = fixtures.make_batch(256, seed=31)
    freq = np.linspace(0.5, 3.0, 250)
    rng = np.random.Generator(np.random.PCG64(7))
    gm = rng.standard_normal((256, 4)).astype(np.float32)
    got = physics.peak_metrics_vjp(spec.to(DEV), torch.from_numpy(gm).to(DEV)).cpu().numpy()
    nan_rows = 0
    for r in range(256):
        idx = int(np.argmin(spec[r].numpy()))
        vals, g = P.peak_parameters_vjp(freq, spec[r].numpy().astype(np.float64), idx, gm[r].astype(np.float64))
        if np.isnan(vals[1]):
            nan_rows += 1
            assert not got[r].any()
            continue
        assert np.count_nonzero(got[r]) <= 5 and set(np.nonzero(got[r])[0]) <= set(np.nonzero(g)[0])
        scale = np.abs(g).max()
        assert np.max(np.abs(got[r] - g)) <= 1e-5 * scale + 1e-12, (r, np.max(np.abs(got[r] - g)), scale)
    assert nan_rows < 64
    # autograd Function: a metric loss on spectra reaches the spectra
    x = spec[:64].to(DEV).clone().requires_grad_(True)
    m = physics.differentiable_peak_metrics(x)
    ok = ~torch.isnan(m[:, 1])
    loss = ((m[ok, 1] - 5.0) ** 2).mean() + 0.1 * m[ok, 2].sum()
    loss.backward()
    assert x.grad is not None and torch.isfinite(x.grad).all() and float(x.grad.abs().sum()) > 0
    assert float(x.grad[~ok].abs().sum()) == 0.0
    # descent direction: a small step along -grad lowers the loss
    with torch.no_grad():
        m2 = physics.differentiable_peak_metrics(x - 1e-3 * x.grad / x.grad.abs().max())
        loss2 = ((m2[ok, 1] - 5.0) ** 2).mean() + 0.1 * m2[ok, 2].sum()
    assert float(loss2) < float(loss.detach())


def test_physics_backward_full_size_properties():
    """4 M spectra: at most five non-zeros per row, zero rows exactly where Q is undefined, forward outputs equal
    the forward kernel's."""
    from pigan_b200 import device_data as DD
    from pigan_b200 import native, physics
    n = 1 << 22
    spec, _ = DD.generate_spectra(n, DEV, seed=4, noise_level=0.1)
    gm = torch.ones(n, 4, device=DEV)
    g = physics.peak_metrics_vjp(spec, gm)
    m = physics.peak_metrics(spec)
    nnz = (g != 0).sum(dim=1)
    assert int(nnz.max()) <= 5
    undefined = torch.isnan(m["Q"])
    assert int(nnz[undefined].max()) == 0 if undefined.any() else True
    assert float((nnz[~undefined] > 0).float().mean()) > 0.999
    assert torch.isfinite(g).all()


def test_two_peak_metrics_match_the_oracle_at_band_minima():
    """physics.two_peak_metrics: band argmins bit-exact against numpy, the eight metric columns against
    oracle/physics.py (calculate_peak_parameters at those indices), 1e-5."""
    from oracle import fixtures
    from oracle import physics as P
    from pigan_b200 import physics
    spec, *_ = fixtures.make_batch(300, seed=41)
    freq = np.linspace(0.5, 3.0, 250)
    k = int((freq < 1.5).sum())
    r = physics.two_peak_metrics(spec.to(DEV))
    i1, i2 = spec[:, :k].numpy().argmin(axis=1), spec[:, k:].numpy().argmin(axis=1) + k
    assert np.array_equal(r["peak_idx"].cpu().numpy(), np.stack([i1, i2], axis=1))
    _, a = P.physics_batch(spec.numpy(), freq, i1.astype(np.int32))
    _, b = P.physics_batch(spec.numpy(), freq, i2.astype(np.int32))
    ref = np.stack([a[:, 0], b[:, 0], a[:, 1], a[:, 2], a[:, 3], b[:, 1], b[:, 2], b[:, 3]], axis=1)
    got = r["metrics"].cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=1e-5)
    assert physics.METRIC_NAMES == ("f1", "f2", "Q1", "FoM1", "S1", "Q2", "FoM2", "S2")
    # the two dips sit where the generator put them
    assert abs(float(np.nanmean(got[:, 0])) - 0.87) < 0.05 and abs(float(np.nanmean(got[:, 1])) - 2.11) < 0.06


def test_synthetic_dataset_carries_physics_metrics_normalised_like_the_reference():
    """DeviceDataset.synthetic: metric columns from the physics kernel at the two band minima, normalised with the
    dataset's rule (min/max over non-NaN, NaN -> 0.5, data_loader.py:198-219)."""
    from pigan_b200 import device_data as DD
    from pigan_b200 import physics
    ds = DD.DeviceDataset.synthetic(4096, DEV, seed=11)
    md, mn = ds.metrics_denorm, ds.metrics_norm
    assert md.shape == (4096, 8) and mn.shape == (4096, 8)
    assert float(mn.min()) >= 0.0 and float(mn.max()) <= 1.0 and not torch.isnan(mn).any()
    again = physics.two_peak_metrics(ds.spectra)["metrics"]
    assert torch.equal(torch.nan_to_num(md, nan=-1.0), torch.nan_to_num(again, nan=-1.0))
    for i, name in enumerate(physics.METRIC_NAMES):
        col = md[:, i]
        ok = ~torch.isnan(col)
        lo, hi = ds.metric_ranges[name]
        assert lo == float(col[ok].min()) and hi == float(col[ok].max())
        torch.testing.assert_close(mn[ok, i], (col[ok] - lo) / (hi - lo), rtol=0, atol=1e-6)
        assert bool((mn[~ok, i] == 0.5).all())
    assert ds.metric_name_to_idx["f2"] == 1 and abs(float(md[:, 0].nanmean()) - 0.87) < 0.05
    assert DD.DeviceDataset.synthetic(64, DEV, seed=1, metrics="uniform").metrics_norm.shape == (64, 8)
