"""ORACLE (test infrastructure): physics metrics of a transmission spectrum.

``peak_parameters`` restates calculate_peak_parameters (reference core/utils/data_loader.py:13-58) as
scalar Python/NumPy float64 for small cases; ``physics_batch`` calls the C restatement (oracle/physics.c)
for millions of spectra.  ``sensitivity`` is the S its callers derive (data_loader.py:96,105).
``peak_parameters_vjp`` is the differentiable restatement used to check ``pigan_physics_metrics_backward``: the
reference has no gradient of these metrics, so that function is pinned to ``peak_parameters`` itself —
same forward values, gradient equal to its central finite differences (tests/test_oracle_golden.py).
Pinned against the reference by tests/golden/physics.npz; only tests/, smoke() and bench.py's CPU legs use it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def peak_parameters(frequency, transmission_db, peak_idx, baseline_transmission=0):
    """data_loader.py:13-58 — returns (f_res, Q, FoM)."""
    f_res = frequency[peak_idx]                                              # :14
    t_min = transmission_db[peak_idx]                                        # :15
    half = t_min + (baseline_transmission - t_min) / 2                       # :16
    f_lower = f_upper = np.nan                                               # :17
    t = transmission_db
    for i in range(peak_idx - 1, -1, -1):                                    # :20
        if (t[i] >= half and t[i + 1] < half) or (t[i] < half and t[i + 1] >= half):
            if (t[i + 1] - t[i]) != 0:
                f_lower = frequency[i] + (half - t[i]) * (frequency[i + 1] - frequency[i]) / (t[i + 1] - t[i])
            else:
                f_lower = frequency[i]
            break
    for i in range(peak_idx + 1, len(frequency) - 1):                        # :32
        if (t[i] <= half and t[i + 1] > half) or (t[i] > half and t[i + 1] <= half):
            if (t[i + 1] - t[i]) != 0:
                f_upper = frequency[i] + (half - t[i]) * (frequency[i + 1] - frequency[i]) / (t[i + 1] - t[i])
            else:
                f_upper = frequency[i]
            break
    q = fom = np.nan                                                         # :43-44
    if not np.isnan(f_lower) and not np.isnan(f_upper) and f_upper > f_lower:  # :47
        delta_f = f_upper - f_lower
        if delta_f > 1e-9:
            q = f_res / delta_f
        if not np.isnan(t_min) and abs(t_min) > 1e-6:                        # :53
            fom = q / abs(t_min) if not np.isnan(q) else np.nan
    return f_res, q, fom


def sensitivity(f, q):
    """data_loader.py:96 / :105."""
    return (f / 1.0) * (q / 100.0) * 100 if not np.isnan(q) else np.nan


def physics_rows_python(spectra: np.ndarray, frequency: np.ndarray, peak_idx=None, baseline=0.0):
    """Row loop over ``peak_parameters`` (small inputs).  Returns (idx int32 [n], metrics float64 [n,4])."""
    n = spectra.shape[0]
    idx = np.empty(n, np.int32)
    out = np.empty((n, 4), np.float64)
    for r in range(n):
        row = spectra[r].astype(np.float64)
        i = int(np.argmin(spectra[r])) if peak_idx is None else int(peak_idx[r])
        f, q, fom = peak_parameters(frequency, row, i, baseline)
        idx[r] = i
        out[r] = (f, q, fom, sensitivity(f, q))
    return idx, out


_lib = None


def _clib():
    global _lib
    if _lib is None:
        from . import build as _b

        _lib = C.CDLL(_b.build())
        _lib.pigan_oracle_physics_batch.restype = None
        _lib.pigan_oracle_physics_batch.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                                    C.c_double, C.c_void_p, C.c_void_p]
    return _lib


def physics_batch(spectra: np.ndarray, frequency: np.ndarray, peak_idx=None, baseline=0.0):
    """C restatement over fp32 spectra [n,s]; returns (idx int32 [n], metrics float64 [n,4])."""
    spectra = np.ascontiguousarray(spectra, np.float32)
    frequency = np.ascontiguousarray(frequency, np.float64)
    n, s = spectra.shape
    assert s <= 4096 and frequency.shape == (s,)
    idx = np.empty(n, np.int32)
    out = np.empty((n, 4), np.float64)
    pk = None
    if peak_idx is not None:
        pk = np.ascontiguousarray(peak_idx, np.int32)
    _clib().pigan_oracle_physics_batch(spectra.ctypes.data, n, s, frequency.ctypes.data,
                                       None if pk is None else pk.ctypes.data, float(baseline),
                                       idx.ctypes.data, out.ctypes.data)
    return idx, out


def peak_parameters_vjp(frequency, transmission_db, peak_idx, grad, baseline_transmission=0):
    """Differentiable restatement of ``peak_parameters`` (+ sensitivity) in torch float64: the branch decisions are
    taken exactly as above (same loops, same comparisons), the arithmetic of :16, :24, :36, :49-54 and :96 is
    replayed on a leaf tensor and autograd returns d(sum_k grad[k] * metric_k)/d(transmission_db).  Returns
    (metrics float64 [4] = f_res, Q, FoM, S with NaN where undefined, gradient float64 [len])."""
    import torch
    f = np.asarray(frequency, dtype=np.float64)
    tn = np.asarray(transmission_db, dtype=np.float64)
    t = torch.tensor(tn, dtype=torch.float64, requires_grad=True)
    f_res = float(f[peak_idx])
    t_min = t[peak_idx]
    half = t_min + (baseline_transmission - t_min) / 2                       # :16
    hv = float(half.detach())
    f_lower = f_upper = None
    for i in range(peak_idx - 1, -1, -1):                                    # :20
        if (tn[i] >= hv and tn[i + 1] < hv) or (tn[i] < hv and tn[i + 1] >= hv):
            if (tn[i + 1] - tn[i]) != 0:
                f_lower = f[i] + (half - t[i]) * (f[i + 1] - f[i]) / (t[i + 1] - t[i])
            else:
                f_lower = torch.tensor(f[i], dtype=torch.float64)
            break
    for i in range(peak_idx + 1, len(f) - 1):                                # :32
        if (tn[i] <= hv and tn[i + 1] > hv) or (tn[i] > hv and tn[i + 1] <= hv):
            if (tn[i + 1] - tn[i]) != 0:
                f_upper = f[i] + (half - t[i]) * (f[i + 1] - f[i]) / (t[i + 1] - t[i])
            else:
                f_upper = torch.tensor(f[i], dtype=torch.float64)
            break
    nan = float("nan")
    q = fom = s = None
    if f_lower is not None and f_upper is not None and float(f_upper.detach()) > float(f_lower.detach()):   # :47
        delta_f = f_upper - f_lower
        if float(delta_f.detach()) > 1e-9:
            q = f_res / delta_f
        if abs(float(t_min.detach())) > 1e-6 and q is not None:              # :53
            fom = q / torch.abs(t_min)
    if q is not None:
        s = (f_res / 1.0) * (q / 100.0) * 100                                # :96
    loss = torch.zeros((), dtype=torch.float64)
    for gk, mk in zip(grad[1:], (q, fom, s)):
        if mk is not None:
            loss = loss + float(gk) * mk
    g = np.zeros_like(tn)
    if loss.requires_grad:
        loss.backward()
        g = t.grad.numpy().copy()
    vals = [f_res] + [float(m.detach()) if m is not None else nan for m in (q, fom, s)]
    return np.array(vals, dtype=np.float64), g
