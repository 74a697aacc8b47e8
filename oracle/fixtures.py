"""ORACLE (test infrastructure): deterministic weights and inputs that do not depend on torch's RNG stream.

numpy's PCG64 bit generator is stable across numpy versions, so the golden-vector generator
(tools/make_golden.py, which runs the real reference) and the tests rebuild bit-identical weights and
batches without shipping multi-megabyte fixtures.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

from . import models as O


def _uniform(rng, shape, bound) -> torch.Tensor:
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def make_weights(seed: int = 42) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """(generator, discriminator, forward_model) state_dicts at the reference widths: Linear weights/biases
    ~ U(-1/sqrt(in), 1/sqrt(in)) like nn.Linear's default, norm layers with non-trivial affine parameters and
    BatchNorm with non-trivial running statistics so every code path is exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g: Dict[str, torch.Tensor] = {}
    for li, (i, o) in zip((0, 3, 6), ((250, 512), (512, 256), (256, 4))):
        b = 1.0 / math.sqrt(i)
        g[f"main.{li}.weight"] = _uniform(rng, (o, i), b)
        g[f"main.{li}.bias"] = _uniform(rng, (o,), b)
    for bi, h in ((1, 512), (4, 256)):
        g[f"main.{bi}.weight"] = _uniform(rng, (h,), 0.5) + 1.0
        g[f"main.{bi}.bias"] = _uniform(rng, (h,), 0.2)
        g[f"main.{bi}.running_mean"] = _uniform(rng, (h,), 0.3)
        g[f"main.{bi}.running_var"] = _uniform(rng, (h,), 0.4) + 1.0
        g[f"main.{bi}.num_batches_tracked"] = torch.tensor(3, dtype=torch.int64)
    g = {k: g[k] for k in ["main.0.weight", "main.0.bias", "main.1.weight", "main.1.bias", "main.1.running_mean",
                           "main.1.running_var", "main.1.num_batches_tracked", "main.3.weight", "main.3.bias",
                           "main.4.weight", "main.4.bias", "main.4.running_mean", "main.4.running_var",
                           "main.4.num_batches_tracked", "main.6.weight", "main.6.bias"]}
    d: Dict[str, torch.Tensor] = {}
    for li, (i, o) in zip((0, 2, 4), ((254, 512), (512, 256), (256, 1))):
        b = 1.0 / math.sqrt(i)
        d[f"main.{li}.weight"] = _uniform(rng, (o, i), b)
        d[f"main.{li}.bias"] = _uniform(rng, (o,), b)
    f: Dict[str, torch.Tensor] = {}
    dims = [4, 256, 512, 1024, 512, 256, 258]
    for k, li in enumerate(O.F_LINEAR):
        i, o = dims[k], dims[k + 1]
        b = 1.0 / math.sqrt(i)
        f[f"model.{li}.weight"] = _uniform(rng, (o, i), b)
        f[f"model.{li}.bias"] = _uniform(rng, (o,), b)
        if k < 5:
            ni = O.F_NORM[k]
            f[f"model.{ni}.weight"] = _uniform(rng, (o,), 0.5) + 1.0
            f[f"model.{ni}.bias"] = _uniform(rng, (o,), 0.2)
    # the last layer's bias places the output near the dB range of real spectra so reconstruction losses are O(1)
    f["model.20.bias"][:250] += -3.0
    return g, d, f


def make_batch(n: int, seed: int = 7, num_points: int = 250, noise_level: float = 0.1):
    """(spectrum [n,S], params_raw [n,4], params_norm [n,4], metrics_norm [n,8]) fp32 — the reference's synthetic
    spectrum formula (core/utils/data_loader.py:62-80) driven by numpy's PCG64."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = np.linspace(0.5, 3.0, num_points)
    p = rng.uniform(2.2, 2.8, size=(n, 4))
    r1, r2, w, g = (p[:, i:i + 1] - 2.5 for i in range(4))
    c1 = 0.870 + 0.05 * r1 + 0.03 * w
    d1 = -12.657 + 1.5 * r2 - 1.0 * g
    w1 = 0.08 + np.abs(0.02 * r1)
    c2 = 2.115 + 0.07 * r2 + 0.04 * g
    d2 = -11.763 + 1.0 * r1 - 0.8 * w
    w2 = 0.15 + np.abs(0.03 * r2)
    t = d1 * np.exp(-((f - c1) ** 2) / (2 * w1 ** 2)) + d2 * np.exp(-((f - c2) ** 2) / (2 * w2 ** 2))
    t = t - 0.5 * (np.tanh((f - 1.5) * 2) + 1) + (-0.5 + 0.5 * (f / 3.0))
    t = t + rng.normal(0, noise_level, size=t.shape)
    t = np.minimum(t, 0)
    spec = torch.from_numpy(t.astype(np.float32))
    praw = torch.from_numpy(p.astype(np.float32))
    pnorm = (praw - 2.2) / 0.6 * 2.0 - 1.0
    mnorm = torch.from_numpy(rng.uniform(0, 1, size=(n, 8)).astype(np.float32))
    return spec, praw, pnorm, mnorm


def sample_indices(numel: int, count: int = 2048) -> torch.Tensor:
    """Fixed strided sample used to pin big tensors (weights, gradients) without storing them whole."""
    if numel <= count:
        return torch.arange(numel)
    return torch.linspace(0, numel - 1, count).long()


F_HIDDEN = (256, 512, 1024, 512, 256)


def make_dropout_masks(n: int, seed: int, p: float = 0.2):
    """Keep-masks (float 0/1) for the five Dropout layers of the forward model, numpy PCG64 — the explicit stand-in
    for torch's RNG stream when the reference's pretraining step is pinned (tools/make_golden.py: fwd_pretrain)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [torch.from_numpy((rng.random((n, h)) >= p).astype(np.float32)) for h in F_HIDDEN]


def evaluator_cases():
    """Seeded (y_true, y_pred) pairs and score arrays of tests/golden/evaluator_metrics.npz (same construction as
    tools/make_golden.py: evaluator_cases / evaluator_metrics)."""
    cases = {}
    spec, praw, pnorm, mnorm = make_batch(300, seed=500)
    rng = np.random.Generator(np.random.PCG64(501))
    cases["spectra"] = (spec.numpy(), (spec + 0.3 * torch.from_numpy(rng.standard_normal(spec.shape).astype(np.float32))).numpy())
    cases["params"] = (praw.numpy(), (praw + 0.05 * torch.from_numpy(rng.standard_normal(praw.shape).astype(np.float32))).numpy())
    cases["metrics"] = (mnorm.numpy(), (0.8 * mnorm + 0.1).numpy())
    y = rng.standard_normal(257).astype(np.float32)
    cases["vector"] = (y, (y + 0.1 * rng.standard_normal(257)).astype(np.float32))
    rng = np.random.Generator(np.random.PCG64(502))
    viol = rng.integers(0, 3, size=1000) * (rng.random(1000) < 0.3)
    err = rng.random(1000).astype(np.float32) * 5
    cons = (1.0 / (1.0 + err)).astype(np.float32)
    return cases, (viol.astype(np.int32), err, cons)


def dataset_csv_columns():
    """Columns of the 32-row synthetic CSV behind tests/golden/dataset.npz (same construction in
    tools/make_golden.py: dataset and tests/test_host_logic.py)."""
    spec, praw, pnorm, _ = make_batch(32, seed=3)
    freqs = np.linspace(0.5, 3.0, 250)
    cols = {f"Freq_{f:.2f}": spec[:, i].numpy() for i, f in enumerate(freqs)}
    for i, n in enumerate(["r1", "r2", "w", "g"]):
        cols[n] = praw[:, i].numpy()
    rng = np.random.default_rng(0)
    metrics = rng.uniform(0.5, 9.0, size=(32, 8))
    metrics[3, 2] = np.nan
    for i, n in enumerate(["f1", "f2", "Q1", "FoM1", "S1", "Q2", "FoM2", "S2"]):
        cols[n] = metrics[:, i]
    return cols
