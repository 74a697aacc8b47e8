"""Drop-in ``core.utils.data_loader``: dataset + normalisation helpers of the reference
(core/utils/data_loader.py:115-330) and the physics-metric helper (:13-58), the latter served by the CUDA
kernel K7 (csrc/physics.cu) — batched, one warp per spectrum.

CSV parsing is host I/O and stays in pandas; the tensors it yields (spectra, raw / normalised parameters and
metrics) are bit-identical to the reference's because the same fp32 operations run in the same order.
"""
import os

import numpy as np
import torch
from torch.utils.data import Dataset

import config.config as cfg  # noqa: F401  (the reference imports it at module import time, :9)

_PARAM_ORDER = ["r1", "r2", "w", "g"]


def batched_peak_parameters(spectra, frequency, peak_idx=None, baseline_transmission=0.0):
    """K7 over a batch: spectra [n,s] (any float dtype/device) -> (peak_idx int32 [n], metrics fp32 [n,4] =
    f_res, Q, FoM, S) on the GPU.  peak_idx=None uses the argmin of every row (first occurrence)."""
    from pigan_b200 import native

    if not torch.cuda.is_available():
        raise RuntimeError("batched_peak_parameters needs a CUDA device (no CPU fallback)")
    dev = spectra.device if isinstance(spectra, torch.Tensor) and spectra.is_cuda else torch.device("cuda")
    spec = torch.as_tensor(spectra).to(dev, torch.float32).contiguous()
    freq = torch.as_tensor(frequency).to(dev, torch.float64).contiguous()
    n, s = spec.shape
    pk = None if peak_idx is None else torch.as_tensor(peak_idx).to(dev, torch.int32).contiguous()
    out_idx = torch.empty(n, device=dev, dtype=torch.int32)
    out = torch.empty(n, 4, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        native.check(native.lib.pigan_physics_metrics(spec.data_ptr(), n, s, freq.data_ptr(), native.ptr(pk),
                                                      float(baseline_transmission), out_idx.data_ptr(),
                                                      out.data_ptr(), native.current_stream()))
    return out_idx, out


def calculate_peak_parameters(frequency, transmission_db, peak_idx, baseline_transmission=0):
    """Reference signature (:13): one spectrum -> (f_res, Q, FoM) as Python floats (NaN where undefined)."""
    spec = np.asarray(transmission_db, dtype=np.float32)[None, :]
    _, m = batched_peak_parameters(spec, np.asarray(frequency, dtype=np.float64), np.asarray([peak_idx]),
                                   baseline_transmission)
    f, q, fom, _ = m[0].tolist()
    return f, q, fom


class MetamaterialDataset(Dataset):
    """Same attributes and 5-tuple items as the reference class (:115-234)."""

    def __init__(self, data_path: str, num_points_per_sample: int = 250, load_data: bool = True):
        self.frequencies = np.linspace(0.5, 3.0, num_points_per_sample)
        self.param_ranges = {name: (2.2, 2.8) for name in _PARAM_ORDER}
        self.metric_names = ["f1", "f2", "Q1", "FoM1", "S1", "Q2", "FoM2", "S2"]
        self.spectrum_cols = [f"Freq_{f:.2f}" for f in self.frequencies]
        self.param_cols = list(_PARAM_ORDER)
        self.metric_cols = self.metric_names
        self.spectra = self.parameters = self.metrics = None
        self.normalized_parameters = self.normalized_metrics = None
        self.metric_ranges = {}
        self.metric_name_to_idx = {n: i for i, n in enumerate(self.metric_names)}
        if load_data:
            self._load(data_path, num_points_per_sample)

    def _load(self, data_path, num_points):
        import pandas as pd

        if not os.path.exists(data_path):
            raise FileNotFoundError(f"data file not found: {data_path}")
        df = pd.read_csv(data_path)

        def is_freq(col):
            parts = col.split("_")
            return col.startswith("Freq_") and len(parts) == 2 and parts[1].replace(".", "", 1).isdigit()

        found = [c for c in df.columns if is_freq(c)]
        if not found:
            raise ValueError("no 'Freq_*' spectrum columns in the CSV")
        self.spectrum_cols = sorted(found, key=lambda c: float(c.split("_")[1]))
        if len(self.spectrum_cols) != num_points:
            print(f"warning: CSV has {len(self.spectrum_cols)} spectrum points, expected {num_points}; using the CSV's")
            self.frequencies = np.linspace(0.5, 3.0, len(self.spectrum_cols))
        missing = [c for c in self.param_cols + self.metric_cols if c not in df.columns]
        if missing:
            raise ValueError(f"CSV is missing required columns: {missing}")
        self.spectra = torch.tensor(df[self.spectrum_cols].values, dtype=torch.float32)
        self.parameters = torch.tensor(df[self.param_cols].values, dtype=torch.float32)
        self.metrics = torch.tensor(df[self.metric_cols].values, dtype=torch.float32)

        # parameters -> [0,1] with the fixed ranges, then -> [-1,1]   (:185-194)
        norm = self.parameters.clone()
        for i, name in enumerate(self.param_cols):
            lo, hi = self.param_ranges[name]
            norm[:, i] = (self.parameters[:, i] - lo) / (hi - lo) if hi - lo > 1e-6 else 0.5
        self.normalized_parameters = norm * 2.0 - 1.0

        # metrics -> [0,1] with data min/max over non-NaN entries, NaN -> 0.5   (:198-219)
        nm = self.metrics.clone()
        for i, name in enumerate(self.metric_names):
            col = self.metrics[:, i]
            ok = col[~torch.isnan(col)]
            lo, hi = (ok.min().item(), ok.max().item()) if len(ok) > 0 else (0.0, 1.0)
            self.metric_ranges[name] = (lo, hi)
            nm[:, i] = (col - lo) / (hi - lo) if hi - lo > 1e-6 else 0.5
        nm[torch.isnan(nm)] = 0.5
        self.normalized_metrics = nm

    def __len__(self):
        return 0 if self.spectra is None else len(self.spectra)

    def __getitem__(self, idx):
        if self.spectra is None:
            raise RuntimeError("dataset not loaded (construct with load_data=True)")
        return (self.spectra[idx], self.parameters[idx], self.normalized_parameters[idx], self.metrics[idx],
                self.normalized_metrics[idx])


def denormalize_params(norm_params_tensor: torch.Tensor, param_ranges: dict) -> torch.Tensor:
    """[-1,1] -> physical range, column by column (:238-252); differentiable."""
    out = torch.zeros_like(norm_params_tensor)
    for i, name in enumerate(_PARAM_ORDER):
        lo, hi = param_ranges[name]
        out[:, i] = (norm_params_tensor[:, i] + 1.0) / 2.0 * (hi - lo) + lo
    return out


def denormalize_metrics(norm_metrics_tensor: torch.Tensor, metric_ranges: dict) -> torch.Tensor:
    """[0,1] -> data range per metric; NaN -> 0 (:255-293)."""
    out = torch.zeros_like(norm_metrics_tensor)
    for i, name in enumerate(list(metric_ranges.keys())):
        lo, hi = metric_ranges[name]
        out[:, i] = norm_metrics_tensor[:, i] * (hi - lo) + lo if hi - lo > 1e-6 else lo
    out[torch.isnan(out)] = 0.0
    return out


def normalize_spectrum(spectrum_tensor: torch.Tensor, global_min_val: float = None,
                       global_max_val: float = None) -> torch.Tensor:
    """Min-max to [0,1] with optional global bounds, clamped (:298-330)."""
    if global_min_val is not None and global_max_val is not None:
        lo, hi = global_min_val, global_max_val
    else:
        lo, hi = spectrum_tensor.min().item(), spectrum_tensor.max().item()
    if hi - lo > 1e-8:
        out = (spectrum_tensor - lo) / (hi - lo)
    else:
        out = torch.full_like(spectrum_tensor, 0.5)
    return torch.clamp(out, 0.0, 1.0)
