"""Generates tests/golden/*.npz by running the REAL reference (/root/reference) on CPU.

Run in the build container only (the reference is not present on the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py
Weights and batches come from oracle/fixtures.py (numpy PCG64), so tests rebuild the same inputs and only
outputs are stored.  What is pinned:
  physics.npz      calculate_peak_parameters on dataset/THZ.txt, on the reference's own synthetic generator
                   (np.random.seed(42)) and on 256 synthetic rows at their argmin
  forward.npz      Generator (eval + train incl. BatchNorm buffer updates), Discriminator, ForwardModel outputs,
                   denormalize_params, the six loss functions
  train_step.npz   core.train.train_pigan.train_pigan itself: 1 epoch x 1 batch (per-step losses, raw gradients
                   captured with hooks) and 3 epochs x 2 batches (schedulers, Adam t>1, BN momentum), sampled
  fwd_pretrain.npz core.train.pretrain_fwd_model.pretrain_forward_model itself: 2 epochs x 2 batches with Dropout
                   fed from explicit masks (epoch losses, first-step raw gradients, final weights, sampled)
  evaluator_metrics.npz  UnifiedEvaluator.calculate_metrics (sklearn/scipy) on seeded arrays + the numpy summary
  datagen.npz      generate_single_terahertz_spectrum_and_params: noise-free rows, rows with (stored) numpy noise
  dataset.npz      MetamaterialDataset on a 32-row synthetic CSV: tensors, item tuple, (de)normalisation helpers
  scoring.npz      the evaluator loop (unified_evaluator.py:369-392) through UnifiedEvaluator itself with the
                   plotting modules stubbed out
"""
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.gridspec", "seaborn"):
    sys.modules[mod] = MagicMock()

import config.config as cfg  # reference config
from core.models.generator import Generator
from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from core.utils import data_loader as ref_dl
from core.utils import loss as ref_loss
from core.train import train_pigan as ref_train

from oracle import fixtures

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(1)  # fixed reduction order


def ref_models(seed=42):
    g_sd, d_sd, f_sd = fixtures.make_weights(seed)
    G = Generator(250, 4); D = Discriminator(250, 4); F = ForwardModel(4, 250, 8)
    G.load_state_dict(g_sd); D.load_state_dict(d_sd); F.load_state_dict(f_sd)
    return G, D, F


def physics():
    rows = [l.split() for l in open(os.path.join(REF, "dataset", "THZ.txt"))
            if l.strip() and l.split()[0].replace(".", "", 1).isdigit()]
    a = np.array(rows, float)
    f, t = a[:, 0], a[:, 1]
    i = int(np.argmin(t))
    thz = ref_dl.calculate_peak_parameters(f, t, i)
    np.random.seed(42)
    freq = np.linspace(0.5, 3.0, 250)
    gen = ref_dl.generate_single_terahertz_spectrum_and_params(freq, 2.5, 2.5, 2.5, 2.5)
    spec, _, _, _ = fixtures.make_batch(256, seed=11)
    spec = spec.numpy()
    idx = np.argmin(spec, axis=1)
    out = np.array([ref_dl.calculate_peak_parameters(freq, spec[r].astype(np.float64), int(idx[r])) for r in range(256)])
    # arbitrary peak indices and a non-zero baseline exercise the other branches
    rng = np.random.Generator(np.random.PCG64(5))
    idx2 = rng.integers(0, 250, size=256)
    out2 = np.array([ref_dl.calculate_peak_parameters(freq, spec[r].astype(np.float64), int(idx2[r]), -0.5)
                     for r in range(256)])
    np.savez(os.path.join(OUT, "physics.npz"), thz_freq=f, thz_t=t, thz_idx=i, thz_out=np.array(thz),
             gen_spectrum=gen[0], gen_metrics=np.array(gen[1:], dtype=np.float64),
             batch_idx=idx.astype(np.int32), batch_out=out, batch_idx2=idx2.astype(np.int32), batch_out2=out2)
    print("physics: THZ", i, thz)


def forward():
    G, D, F = ref_models()
    spec, praw, pnorm, mnorm = fixtures.make_batch(64, seed=7)
    res = {}
    with torch.no_grad():
        G.eval(); res["g_eval"] = G(spec).numpy()
        G.train(); res["g_train"] = G(spec).numpy()
        for k in ("main.1.running_mean", "main.1.running_var", "main.4.running_mean", "main.4.running_var",
                  "main.1.num_batches_tracked"):
            res["g_after_" + k] = G.state_dict()[k].numpy()
        res["d_out"] = D(spec, praw).numpy()
        F.eval()
        fs, fm = F(pnorm)
        res["f_spec"], res["f_metrics"] = fs.numpy(), fm.numpy()
        ds = ref_dl.MetamaterialDataset("", load_data=False)
        res["denorm"] = ref_dl.denormalize_params(pnorm, ds.param_ranges).numpy()
        res["loss_maxwell"] = ref_loss.maxwell_equation_loss(fs, None, pnorm).numpy()
        res["loss_lc"] = ref_loss.lc_model_approx_loss(fm[:, 0:1], fm[:, 1:2], pnorm).numpy()
        res["loss_range"] = ref_loss.structural_param_range_loss(pnorm * 1.3).numpy()
        res["loss_bce"] = ref_loss.criterion_bce()(D(spec, praw), torch.full((64, 1), 0.9)).numpy()
        res["loss_mse"] = ref_loss.criterion_mse()(fs, spec).numpy()
    np.savez(os.path.join(OUT, "forward.npz"), **res)
    print("forward: g_eval[0]", res["g_eval"][0])


def train_step():
    res = {}
    tmp = tempfile.mkdtemp()
    cfg.CHECKPOINT_DIR = os.path.join(tmp, "ckpt")
    cfg.SAVED_MODELS_DIR = os.path.join(tmp, "saved")
    ds = ref_dl.MetamaterialDataset("", load_data=False)

    def batches(n_batches, B, seed0):
        out = []
        for i in range(n_batches):
            spec, praw, pnorm, mnorm = fixtures.make_batch(B, seed=seed0 + i)
            out.append((spec, praw, pnorm, torch.zeros(B, 8), mnorm))
        return out

    # (a) one epoch, one batch: exact per-step losses + raw (unclipped) gradients through hooks
    G, D, F = ref_models()
    grads = {"g": {}, "d": {}}
    for tag, net in (("g", G), ("d", D)):
        for name, p in net.named_parameters():
            def hook(gr, tag=tag, name=name):
                if name not in grads[tag]:          # first call = D-step for D params, G-step for G params
                    grads[tag][name] = gr.detach().clone()
                return None
            p.register_hook(hook)
    hist = ref_train.train_pigan(batches(1, 64, 100), torch.device("cpu"), G, D, F, ds, num_epochs=1, log_interval=10)
    for k, v in hist.items():
        res["a_" + k] = np.array(v, dtype=np.float64)
    for tag in ("g", "d"):
        for name, gr in grads[tag].items():
            res[f"a_grad_{tag}_{name}"] = gr.reshape(-1)[fixtures.sample_indices(gr.numel())].numpy()
            res[f"a_gradnorm_{tag}_{name}"] = np.array(gr.norm().item())
    for tag, net in (("g", G), ("d", D)):
        for name, t in net.state_dict().items():
            tt = t.reshape(-1).double() if t.is_floating_point() else t.reshape(-1)
            res[f"a_final_{tag}_{name}"] = tt[fixtures.sample_indices(tt.numel())].numpy()

    # (b) three epochs, two batches each: LR schedulers, Adam bias correction, BN momentum
    G, D, F = ref_models()
    hist = ref_train.train_pigan(batches(2, 64, 200), torch.device("cpu"), G, D, F, ds, num_epochs=3, log_interval=10)
    for k, v in hist.items():
        res["b_" + k] = np.array(v, dtype=np.float64)
    for tag, net in (("g", G), ("d", D)):
        for name, t in net.state_dict().items():
            tt = t.reshape(-1).double() if t.is_floating_point() else t.reshape(-1)
            res[f"b_final_{tag}_{name}"] = tt[fixtures.sample_indices(tt.numel())].numpy()
    np.savez(os.path.join(OUT, "train_step.npz"), **res)
    print("train_step: a losses", {k: v for k, v in res.items() if k.startswith("a_") and k.endswith("losses")})


def scoring():
    from torch.utils.data import Dataset
    import core.evaluate.unified_evaluator as ue

    G, D, F = ref_models()
    G.eval(); F.eval()
    spec, praw, pnorm, mnorm = fixtures.make_batch(96, seed=31)

    class DS(Dataset):
        param_ranges = {k: (2.2, 2.8) for k in ("r1", "r2", "w", "g")}
        def __len__(self): return spec.shape[0]
        def __getitem__(self, i): return spec[i], praw[i], pnorm[i], torch.zeros(8), mnorm[i]

    ev = ue.UnifiedEvaluator.__new__(ue.UnifiedEvaluator)
    ev.device = torch.device("cpu"); ev.generator = G; ev.forward_model = F; ev.discriminator = D; ev.dataset = DS()
    np.random.seed(0)
    agg = ev.evaluate_structural_prediction(num_samples=96)   # all 96 rows, permuted
    res = {"agg_" + k: np.array(v, dtype=np.float64) for k, v in agg.items()}
    with torch.no_grad():  # per-row values of the same loop body (:376-392)
        p = G(spec)
        res["params"] = p.numpy()
        res["violations"] = torch.sum((p < 0) | (p > 1), dim=1).numpy()
        rec, _ = F(p)
        err = torch.mean((spec - rec) ** 2, dim=1)
        res["recon_error"] = err.numpy()
        res["consistency"] = (1.0 / (1.0 + err)).numpy()
    np.savez(os.path.join(OUT, "scoring.npz"), **res)
    print("scoring:", {k: float(v) for k, v in agg.items()})


def validation():
    """The model-validation loop (unified_evaluator.py:415-490) through UnifiedEvaluator.evaluate_model_validation
    itself: cycle-consistency error, prediction stability under 0.01 * randn noise, plausibility score.  torch's
    randn_like is replaced by a recorded noise source for the duration of the call (batch size 64, rows in the order
    the reference's np.random.choice picked them), so the noise travels with the golden file."""
    from torch.utils.data import Dataset
    import core.evaluate.unified_evaluator as ue

    G, D, F = ref_models()
    G.eval(); F.eval()
    n = 96
    spec, praw, pnorm, mnorm = fixtures.make_batch(n, seed=32)

    class DS(Dataset):
        param_ranges = {k: (2.2, 2.8) for k in ("r1", "r2", "w", "g")}
        def __len__(self): return n
        def __getitem__(self, i): return spec[i], praw[i], pnorm[i], torch.zeros(8), mnorm[i]

    ev = ue.UnifiedEvaluator.__new__(ue.UnifiedEvaluator)
    ev.device = torch.device("cpu"); ev.generator = G; ev.forward_model = F; ev.discriminator = D; ev.dataset = DS()
    gen = torch.Generator().manual_seed(77)
    drawn = []
    orig = torch.randn_like

    def recorded(t, **kw):
        z = torch.randn(t.shape, generator=gen, dtype=t.dtype)
        drawn.append(z.clone())
        return z

    np.random.seed(1)
    order = np.random.choice(n, n, replace=False)     # the rows evaluate_model_validation will pick
    np.random.seed(1)
    torch.randn_like = recorded
    try:
        agg = ev.evaluate_model_validation(num_samples=n)
    finally:
        torch.randn_like = orig
    noise_in_order = torch.cat(drawn)                  # row j of the loop = dataset row order[j]
    noise = torch.empty_like(noise_in_order)
    noise[torch.as_tensor(order)] = noise_in_order
    res = {"agg_" + k: np.array(v, dtype=np.float64) for k, v in agg.items()}
    res["noise"] = noise.numpy()
    with torch.no_grad():                              # per-row values of the same loop body (:446-468)
        p = G(spec)
        rec, _ = F(p)
        res["cycle_error"] = torch.mean((spec - rec) ** 2, dim=1).numpy()
        pn_ = G(spec + noise * 0.01)
        res["stability"] = torch.mean((p - pn_) ** 2, dim=1).numpy()
        res["plausibility"] = torch.mean(torch.sigmoid(p * 10 - 5), dim=1).numpy()
    for k in ("cycle_error", "stability", "plausibility"):   # the per-row restatement reproduces the aggregates
        key = {"cycle_error": "cycle_consistency_error", "stability": "prediction_stability",
               "plausibility": "physical_plausibility"}[k]
        assert abs(res[k].mean() - agg[key + "_mean"]) <= 1e-6 * abs(agg[key + "_mean"]) + 1e-12, (k, res[k].mean(), agg)
    np.savez(os.path.join(OUT, "validation.npz"), **res)
    print("validation:", {k: float(v) for k, v in agg.items()})


def fwd_pretrain():
    """The reference's own pretrain_forward_model (pretrain_fwd_model.py:24-158) with torch.nn.functional.dropout
    replaced by explicit keep-masks (oracle/fixtures.make_dropout_masks) so the run is reproducible elsewhere."""
    import torch.nn.functional as TF
    from core.train import pretrain_fwd_model as ref_pre
    res = {}
    tmp = tempfile.mkdtemp()
    cfg.SAVED_MODELS_DIR = os.path.join(tmp, "saved")
    B, n_batches, epochs, lr = 64, 2, 2, 1e-3
    data = []
    for i in range(n_batches):
        spec, praw, pnorm, mnorm = fixtures.make_batch(B, seed=300 + i)
        data.append((spec, praw, pnorm, torch.zeros(B, 8), mnorm))
    state = {"call": 0}

    def fake_dropout(x, p=0.5, training=True, inplace=False):
        assert training and abs(p - 0.2) < 1e-12
        step, layer = divmod(state["call"], 5)
        state["call"] += 1
        m = fixtures.make_dropout_masks(B, seed=1000 + step)[layer]
        assert m.shape == x.shape
        return x * m / (1.0 - p)

    _, _, F = ref_models()
    first = {}
    for name, prm in F.named_parameters():
        def hook(gr, name=name):
            if name not in first:
                first[name] = gr.detach().clone()
            return None
        prm.register_hook(hook)
    real = TF.dropout
    TF.dropout = fake_dropout
    try:
        hist = ref_pre.pretrain_forward_model(F, data, torch.device("cpu"), num_epochs=epochs, lr=lr, log_interval=10)
    finally:
        TF.dropout = real
    assert state["call"] == 5 * n_batches * epochs
    res["epoch_losses"] = np.array(hist, dtype=np.float64)
    for name, gr in first.items():
        res[f"grad_{name}"] = gr.reshape(-1)[fixtures.sample_indices(gr.numel())].numpy()
        res[f"gradnorm_{name}"] = np.array(gr.norm().item())
    for name, t in F.state_dict().items():
        tt = t.reshape(-1).double()
        res[f"final_{name}"] = tt[fixtures.sample_indices(tt.numel())].numpy()
    np.savez(os.path.join(OUT, "fwd_pretrain.npz"), **res)
    print("fwd_pretrain: epoch losses", hist)


def datagen():
    """generate_single_terahertz_spectrum_and_params itself (data_loader.py:62-111): noise-free rows for seeded
    parameters, and rows with numpy noise (the noise is recovered as output - noise-free output where unclamped,
    so it is stored explicitly by re-drawing it from the same seed)."""
    res = {}
    freq = np.linspace(0.5, 3.0, 250)
    rng = np.random.Generator(np.random.PCG64(600))
    params = 2.2 + 0.6 * rng.random((24, 4))
    params[0] = 2.5                      # |.| kinks of the width terms
    params[1] = (2.2, 2.8, 2.2, 2.8)
    clean, clean_nooff, noisy, noise = [], [], [], []
    for i, (r1, r2, w, g) in enumerate(params):
        clean.append(ref_dl.generate_single_terahertz_spectrum_and_params(freq, r1, r2, w, g, noise_level=0.0)[0])
        clean_nooff.append(ref_dl.generate_single_terahertz_spectrum_and_params(freq, r1, r2, w, g, apply_offset=False,
                                                                                 noise_level=0.0)[0])
        np.random.seed(700 + i)
        noisy.append(ref_dl.generate_single_terahertz_spectrum_and_params(freq, r1, r2, w, g, noise_level=0.1)[0])
        np.random.seed(700 + i)
        noise.append(np.random.normal(0, 0.1, 250) / 0.1)
    res["params"] = params
    res["clean"] = np.array(clean); res["clean_nooffset"] = np.array(clean_nooff)
    res["noisy"] = np.array(noisy); res["noise_unit"] = np.array(noise)
    np.savez(os.path.join(OUT, "datagen.npz"), **res)
    print("datagen:", res["clean"].shape, float(res["clean"].min()))


def dataset():
    """The reference's MetamaterialDataset (data_loader.py:115-234) on a 32-row synthetic CSV, its item tuple and
    the normalisation helpers (:238-330)."""
    import pandas as pd
    res = {}
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "data.csv")
    pd.DataFrame(fixtures.dataset_csv_columns()).to_csv(path, index=False)
    ds = ref_dl.MetamaterialDataset(path, num_points_per_sample=250)
    for name in ("spectra", "parameters", "metrics", "normalized_parameters", "normalized_metrics"):
        res[name] = np.asarray(getattr(ds, name))
    res["frequencies"] = np.asarray(ds.frequencies)
    for k, (lo, hi) in ds.metric_ranges.items():
        res[f"range_{k}"] = np.array([lo, hi], dtype=np.float64)
    item = ds[5]
    for i, t in enumerate(item):
        res[f"item5_{i}"] = t.numpy()
    res["denorm_params"] = ref_dl.denormalize_params(torch.as_tensor(ds.normalized_parameters), ds.param_ranges).numpy()
    res["denorm_metrics"] = ref_dl.denormalize_metrics(torch.as_tensor(ds.normalized_metrics), ds.metric_ranges).numpy()
    res["norm_spectrum"] = ref_dl.normalize_spectrum(torch.as_tensor(ds.spectra)).numpy()
    res["len"] = np.array(len(ds))
    np.savez(os.path.join(OUT, "dataset.npz"), **res)
    print("dataset:", {k: v.shape for k, v in res.items() if v.ndim == 2})


def config_constants():
    """Scalar constants of the reference's config/config.py (paths and the device string excluded)."""
    import json
    ref = {n: getattr(cfg, n) for n in dir(cfg)
           if n.isupper() and isinstance(getattr(cfg, n), (int, float, str))
           and not n.endswith(("_DIR", "_PATH", "ROOT")) and n != "DEVICE"}
    json.dump(ref, open(os.path.join(OUT, "config.json"), "w"), indent=1, sort_keys=True)
    print("config:", len(ref), "constants")


def _structure(obj):
    """JSON-able description of a checkpoint object: dict keys, tensor shapes / dtypes, scalar types."""
    if isinstance(obj, dict):
        return {str(k): _structure(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_structure(v) for v in obj]
    if isinstance(obj, torch.Tensor):
        return {"tensor": list(obj.shape), "dtype": str(obj.dtype)}
    if isinstance(obj, bool) or obj is None:
        return {"scalar": type(obj).__name__}
    if isinstance(obj, (int, float)):
        return {"scalar": type(obj).__name__}
    return {"other": type(obj).__name__}


def checkpoint_structure():
    """Files train_pigan writes (train_pigan.py:284-309) and the layout of their contents: one epoch with
    SAVE_MODEL_INTERVAL = 1."""
    import json
    tmp = tempfile.mkdtemp()
    cfg.CHECKPOINT_DIR = os.path.join(tmp, "ckpt")
    cfg.SAVED_MODELS_DIR = os.path.join(tmp, "saved")
    old = cfg.SAVE_MODEL_INTERVAL
    cfg.SAVE_MODEL_INTERVAL = 1
    try:
        ds = ref_dl.MetamaterialDataset("", load_data=False)
        G, D, F = ref_models()
        spec, praw, pnorm, mnorm = fixtures.make_batch(64, seed=100)
        ref_train.train_pigan([(spec, praw, pnorm, torch.zeros(64, 8), mnorm)], torch.device("cpu"), G, D, F, ds,
                              num_epochs=1, log_interval=10)
    finally:
        cfg.SAVE_MODEL_INTERVAL = old
    res = {"checkpoint_files": sorted(os.listdir(cfg.CHECKPOINT_DIR)), "saved_files": sorted(os.listdir(cfg.SAVED_MODELS_DIR))}
    ck = torch.load(os.path.join(cfg.CHECKPOINT_DIR, "pigan_epoch_1.pth"), weights_only=False)
    res["checkpoint"] = _structure(ck)
    res["loss_history"] = _structure(torch.load(os.path.join(cfg.SAVED_MODELS_DIR, "pigan_loss_history.pt"),
                                                weights_only=False))
    res["param_group_g"] = {k: v for k, v in ck["optimizer_g_state_dict"]["param_groups"][0].items()
                            if isinstance(v, (int, float, bool, type(None))) or k in ("betas", "params")}
    json.dump(res, open(os.path.join(OUT, "checkpoint_structure.json"), "w"), indent=1, sort_keys=True, default=list)
    print("checkpoint_structure:", res["checkpoint_files"], res["saved_files"], list(ck.keys()))


def evaluator_cases():
    """Seeded inputs of the evaluator-reduction golden (rebuilt identically by the tests)."""
    cases = {}
    spec, praw, pnorm, mnorm = fixtures.make_batch(300, seed=500)
    rng = np.random.Generator(np.random.PCG64(501))
    cases["spectra"] = (spec.numpy(), (spec + 0.3 * torch.from_numpy(rng.standard_normal(spec.shape).astype(np.float32))).numpy())
    cases["params"] = (praw.numpy(), (praw + 0.05 * torch.from_numpy(rng.standard_normal(praw.shape).astype(np.float32))).numpy())
    cases["metrics"] = (mnorm.numpy(), (0.8 * mnorm + 0.1).numpy())
    y = rng.standard_normal(257).astype(np.float32)
    cases["vector"] = (y, (y + 0.1 * rng.standard_normal(257)).astype(np.float32))
    return cases


def evaluator_metrics():
    """UnifiedEvaluator.calculate_metrics itself (sklearn + scipy; unified_evaluator.py:138-184) and the numpy
    summary of the structural-prediction loop (:393-405) on seeded arrays."""
    from core.evaluate.unified_evaluator import UnifiedEvaluator
    res = {}
    for name, (y, p) in evaluator_cases().items():
        m = UnifiedEvaluator.calculate_metrics(None, y, p)
        for k, v in m.items():
            res[f"{name}_{k}"] = np.array(float(v))
    rng = np.random.Generator(np.random.PCG64(502))
    viol = rng.integers(0, 3, size=1000) * (rng.random(1000) < 0.3)
    err = rng.random(1000).astype(np.float32) * 5
    cons = (1.0 / (1.0 + err)).astype(np.float32)
    res["summ_param_range_violation_rate"] = np.array(np.mean(viol > 0))
    res["summ_avg_param_violations"] = np.array(np.mean(viol))
    res["summ_reconstruction_error_mean"] = np.array(float(np.mean(err)))
    res["summ_reconstruction_error_std"] = np.array(float(np.std(err)))
    res["summ_consistency_score_mean"] = np.array(float(np.mean(cons)))
    res["summ_consistency_score_std"] = np.array(float(np.std(cons)))
    np.savez(os.path.join(OUT, "evaluator_metrics.npz"), **res)
    print("evaluator_metrics:", {k: float(v) for k, v in res.items() if k.startswith("spectra")})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()[name]()
    else:
        physics(); forward(); train_step(); scoring(); validation(); fwd_pretrain(); evaluator_metrics(); datagen(); dataset(); config_constants(); checkpoint_structure()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
