"""Physics metrics of spectra on the device (csrc/physics.cu through ``pigan_physics_metrics``): the quantities the
north star's PINN loss is built from — resonance peak, FWHM-based Q, FoM, sensitivity S
(``calculate_peak_parameters``, core/utils/data_loader.py:13-58, and its callers :96,105) — plus the peak shift
between two spectra.  The reference defines no peak shift; it is taken as in SURVEY 8(f) N2:
``peak_shift = f_res(reconstructed) - f_res(target)`` with f_res at each row's argmin."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import native, synthetic


def peak_metrics(spectra: torch.Tensor, frequency: Optional[torch.Tensor] = None,
                 peak_idx: Optional[torch.Tensor] = None, baseline_transmission: float = 0.0
                 ) -> Dict[str, torch.Tensor]:
    """spectra [n,S] fp32 CUDA -> {'peak_idx' int32 [n], 'f_res', 'Q', 'FoM', 'S' fp32 [n]} (NaN where the reference
    returns NaN).  ``peak_idx=None``: argmin of every row, first occurrence."""
    if not spectra.is_cuda:
        raise RuntimeError("peak_metrics needs CUDA tensors — the B200 path has no CPU fallback")
    spec = spectra.float().contiguous()
    n, s = spec.shape
    freq = synthetic.frequencies(s, device=spec.device) if frequency is None \
        else torch.as_tensor(frequency).to(spec.device, torch.float64).contiguous()
    pk = None if peak_idx is None else peak_idx.to(spec.device, torch.int32).contiguous()
    idx = torch.empty(n, device=spec.device, dtype=torch.int32)
    out = torch.empty(n, 4, device=spec.device, dtype=torch.float32)
    native.check(native.lib.pigan_physics_metrics(spec.data_ptr(), n, s, freq.data_ptr(), native.ptr(pk),
                                                  float(baseline_transmission), idx.data_ptr(), out.data_ptr(),
                                                  native.current_stream()))
    return {"peak_idx": idx, "f_res": out[:, 0], "Q": out[:, 1], "FoM": out[:, 2], "S": out[:, 3]}


def peak_shift(reconstructed: torch.Tensor, target: torch.Tensor, frequency: Optional[torch.Tensor] = None
               ) -> Dict[str, torch.Tensor]:
    """Per row: resonance frequency of ``reconstructed`` minus that of ``target`` (THz), with both peak indices."""
    a, b = peak_metrics(reconstructed, frequency), peak_metrics(target, frequency)
    return {"peak_shift": a["f_res"] - b["f_res"], "peak_idx_reconstructed": a["peak_idx"],
            "peak_idx_target": b["peak_idx"]}


def peak_metrics_vjp(spectra: torch.Tensor, grad_metrics: torch.Tensor, frequency: Optional[torch.Tensor] = None,
                     peak_idx: Optional[torch.Tensor] = None, baseline_transmission: float = 0.0) -> torch.Tensor:
    """dL/d(spectra) [n,S] from dL/d(f_res, Q, FoM, S) [n,4] (``pigan_physics_metrics_backward``): at most five
    non-zeros per row, zero where Q is undefined."""
    if not spectra.is_cuda:
        raise RuntimeError("peak_metrics_vjp needs CUDA tensors — the B200 path has no CPU fallback")
    spec = spectra.detach().float().contiguous()
    n, s = spec.shape
    gm = grad_metrics.detach().to(spec.device, torch.float32).contiguous()
    if gm.shape != (n, 4):
        raise ValueError(f"grad_metrics must be [{n}, 4]")
    freq = synthetic.frequencies(s, device=spec.device) if frequency is None \
        else torch.as_tensor(frequency).to(spec.device, torch.float64).contiguous()
    pk = None if peak_idx is None else peak_idx.to(spec.device, torch.int32).contiguous()
    out = torch.empty_like(spec)
    native.check(native.lib.pigan_physics_metrics_backward(spec.data_ptr(), n, s, freq.data_ptr(), native.ptr(pk),
                                                           float(baseline_transmission), gm.data_ptr(), out.data_ptr(),
                                                           None, None, native.current_stream()))
    return out


class _PeakMetrics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spectra, frequency, baseline):
        m = peak_metrics(spectra, frequency, None, baseline)
        ctx.save_for_backward(spectra)
        ctx.frequency, ctx.baseline, ctx.peak_idx = frequency, baseline, m["peak_idx"]
        return torch.stack([m["f_res"], m["Q"], m["FoM"], m["S"]], dim=1)

    @staticmethod
    def backward(ctx, grad_out):
        (spectra,) = ctx.saved_tensors
        # NaN outputs (no half-depth crossing) carry no gradient; incoming NaNs must not poison the rows that have one
        g = torch.nan_to_num(grad_out, nan=0.0)
        return peak_metrics_vjp(spectra, g, ctx.frequency, ctx.peak_idx, ctx.baseline), None, None


def differentiable_peak_metrics(spectra: torch.Tensor, frequency: Optional[torch.Tensor] = None,
                                baseline_transmission: float = 0.0) -> torch.Tensor:
    """[n,4] = (f_res, Q, FoM, S) of every spectrum with a backward pass into the spectra, e.g. a physics-metric loss
    ``((m - m_target)[ok] ** 2).mean()`` on a surrogate's predicted spectra (rows where Q is NaN: mask them out)."""
    return _PeakMetrics.apply(spectra, frequency, float(baseline_transmission))


METRIC_NAMES = ("f1", "f2", "Q1", "FoM1", "S1", "Q2", "FoM2", "S2")   # data_loader.py:131


def two_peak_metrics(spectra: torch.Tensor, frequency: Optional[torch.Tensor] = None, split_thz: float = 1.5
                     ) -> Dict[str, torch.Tensor]:
    """The dataset's eight metric columns (f1, f2, Q1, FoM1, S1, Q2, FoM2, S2) from the spectra themselves: the band
    below ``split_thz`` holds the first dip of the reference's generator (centre ~0.87 THz, data_loader.py:64), the
    band above it the second (~2.1 THz, :69).  Each band's resonance index is the argmin inside the band (first
    occurrence); Q / FoM / S come from ``calculate_peak_parameters`` at that index over the WHOLE spectrum, as the
    reference evaluates it (:95, :104).  Returns the [n,8] tensor ``metrics`` in the reference's column order plus
    ``peak_idx`` [n,2].  (SURVEY 8(f) N2's two-peak variant; the reference picks its peaks with scipy.find_peaks.)"""
    if not spectra.is_cuda:
        raise RuntimeError("two_peak_metrics needs CUDA tensors — the B200 path has no CPU fallback")
    spec = spectra.float().contiguous()
    s = spec.shape[1]
    freq = synthetic.frequencies(s, device=spec.device) if frequency is None \
        else torch.as_tensor(frequency).to(spec.device, torch.float64).contiguous()
    k = int((freq < split_thz).sum().item())
    if k < 1 or k >= s:
        raise ValueError("split_thz must fall inside the frequency grid")
    i1 = spec[:, :k].argmin(dim=1)
    i2 = spec[:, k:].argmin(dim=1) + k
    m1, m2 = peak_metrics(spec, freq, i1), peak_metrics(spec, freq, i2)
    cols = [m1["f_res"], m2["f_res"], m1["Q"], m1["FoM"], m1["S"], m2["Q"], m2["FoM"], m2["S"]]
    return {"metrics": torch.stack(cols, dim=1), "peak_idx": torch.stack([i1, i2], dim=1).to(torch.int32)}


# ------------------------------------------------------------------------------------------ multi-GPU
def shard_rows(n: int, rank: int, world: int):
    """Contiguous row range [begin, end) of ``rank`` when n spectra are split over ``world`` GPUs: the first
    ``n % world`` ranks take one row more.  The physics kernel treats every spectrum independently, so the sharded
    result is the concatenation of the ranks' results - no collective on the data path."""
    if not (0 <= rank < world):
        raise ValueError("0 <= rank < world required")
    base, extra = divmod(int(n), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def sharded_peak_summary(local: Dict[str, torch.Tensor], group=None) -> Dict[str, float]:
    """Global summary of a sharded ``peak_metrics`` run: number of spectra, number with a defined Q, and the means of
    f_res / Q / FoM over the defined ones.  The only collective is ONE all-reduce of seven float64 sums; the per-row
    results stay on the rank that computed them."""
    import torch.distributed as dist
    q = local["Q"]
    ok = ~torch.isnan(q)
    sums = torch.stack([torch.tensor(float(q.numel()), device=q.device, dtype=torch.float64),
                        ok.sum().double(),
                        local["f_res"].double().sum(),
                        torch.where(ok, q, torch.zeros_like(q)).double().sum(),
                        torch.where(ok, local["FoM"], torch.zeros_like(q)).nan_to_num(0.0).double().sum(),
                        (~torch.isnan(local["FoM"])).sum().double(),
                        torch.where(ok, local["S"], torch.zeros_like(q)).double().sum()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, group=group)
    n, nq, sf, sq, sfom, nfom, ss = (float(x) for x in sums.tolist())
    return {"spectra": n, "defined_Q": nq, "f_res_mean": sf / max(n, 1.0), "Q_mean": sq / max(nq, 1.0),
            "FoM_mean": sfom / max(nfom, 1.0), "S_mean": ss / max(nq, 1.0)}
