"""Synthetic THz spectra of the dataset's shape, generated directly on the target device.

Vectorised form of the reference's generator formula (core/utils/data_loader.py:62-80: two Gaussian dips
whose centre/depth/width depend on r1,r2,w,g, a tanh step, a linear offset, N(0, noise) noise, clipped at
0 dB) — used for benchmarks, smoke tests and parity inputs because the real CSV is not distributed.
"""
from __future__ import annotations

import torch


def frequencies(num_points: int = 250, device="cpu", dtype=torch.float64) -> torch.Tensor:
    # data_loader.py:124  np.linspace(0.5, 3.0, num_points)
    return torch.linspace(0.5, 3.0, num_points, dtype=torch.float64, device=device).to(dtype)


def make_batch(n: int, num_points: int = 250, seed: int = 42, device="cpu", noise_level: float = 0.1,
               chunk: int = 1 << 20):
    """Returns (spectrum [n,S] fp32, params_raw [n,4] fp32 in (2.2,2.8), params_norm [n,4] in (-1,1),
    metrics_norm [n,8] fp32 in (0,1)).  Deterministic per (seed, device type)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f = torch.linspace(0.5, 3.0, num_points, dtype=torch.float32, device=device)
    spec = torch.empty(n, num_points, dtype=torch.float32, device=device)
    params = torch.rand(n, 4, generator=g, device=device, dtype=torch.float32) * 0.6 + 2.2
    for s in range(0, n, chunk):
        p = params[s:s + chunk]
        r1, r2, w, gg = (p[:, i:i + 1] - 2.5 for i in range(4))
        c1 = 0.870 + 0.05 * r1 + 0.03 * w
        d1 = -12.657 + 1.5 * r2 - 1.0 * gg
        w1 = 0.08 + (0.02 * r1).abs()
        c2 = 2.115 + 0.07 * r2 + 0.04 * gg
        d2 = -11.763 + 1.0 * r1 - 0.8 * w
        w2 = 0.15 + (0.03 * r2).abs()
        t = d1 * torch.exp(-((f - c1) ** 2) / (2 * w1 ** 2))
        t = t + d2 * torch.exp(-((f - c2) ** 2) / (2 * w2 ** 2))
        t = t - 0.5 * (torch.tanh((f - 1.5) * 2) + 1)
        t = t + (-0.5 + 0.5 * (f / 3.0))
        t = t + noise_level * torch.randn(t.shape, generator=g, device=device, dtype=torch.float32)
        spec[s:s + chunk] = torch.clamp(t, max=0.0)
    params_norm = (params - 2.2) / 0.6 * 2.0 - 1.0   # data_loader.py:185-194
    metrics_norm = torch.rand(n, 8, generator=g, device=device, dtype=torch.float32)
    return spec, params, params_norm, metrics_norm
