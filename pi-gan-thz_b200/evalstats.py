"""Device-side evaluator reductions: the numbers UnifiedEvaluator derives from whole result arrays
(core/evaluate/unified_evaluator.py:138-184 ``calculate_metrics``; :393-405 summary of the structural-prediction loop)
without the per-batch ``.cpu().numpy()`` round trips.  Batches accumulate into fp64 sums on the device; with a process
group the sums are all-reduced before the final formulas, so every rank reports the metrics of the whole job."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import native
from .native import check, lib

REGRESSION_KEYS = ("mse", "mae", "rmse", "r2", "pearson_r", "mape")
SUMMARY_KEYS = ("param_range_violation_rate", "avg_param_violations", "reconstruction_error_mean",
                "reconstruction_error_std", "consistency_score_mean", "consistency_score_std")


def _f32(x: torch.Tensor, name: str) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor — the B200 path has no CPU fallback")
    return x.float().contiguous()


class RegressionMetrics:
    """Streaming version of ``calculate_metrics(y_true, y_pred)``: ``update`` per batch, ``compute`` once."""

    def __init__(self, cols: int, device, process_group=None):
        self.cols, self.device, self.pg = int(cols), torch.device(device), process_group
        self.sums = torch.zeros(self.cols * 8, device=self.device, dtype=torch.float64)
        self.ws = torch.empty(lib.pigan_eval_workspace_bytes(self.cols), device=self.device, dtype=torch.uint8)
        self.n = 0

    def update(self, y_true: torch.Tensor, y_pred: torch.Tensor) -> None:
        y, p = _f32(y_true, "y_true"), _f32(y_pred, "y_pred")
        if y.dim() == 1:
            y, p = y[:, None], p[:, None]
        if y.shape != p.shape or y.shape[1] != self.cols:
            raise ValueError(f"expected two [n, {self.cols}] tensors, got {tuple(y.shape)} and {tuple(p.shape)}")
        check(lib.pigan_regression_sums(y.data_ptr(), p.data_ptr(), y.shape[0], self.cols, self.sums.data_ptr(), 1,
                                        self.ws.data_ptr(), self.ws.numel(), native.current_stream()))
        self.n += y.shape[0]

    def compute(self) -> Dict[str, float]:
        sums, n = self.sums, self.n
        if self.pg is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            sums = sums.clone()
            cnt = torch.tensor([n], device=self.device, dtype=torch.int64)
            dist.all_reduce(sums, group=self.pg)
            dist.all_reduce(cnt, group=self.pg)
            n = int(cnt.item())
        out = torch.empty(6, device=self.device, dtype=torch.float64)
        check(lib.pigan_regression_finalize(sums.data_ptr(), n, self.cols, out.data_ptr(), native.current_stream()))
        return dict(zip(REGRESSION_KEYS, out.cpu().tolist()))


def regression_metrics(y_true: torch.Tensor, y_pred: torch.Tensor, process_group=None) -> Dict[str, float]:
    cols = 1 if y_true.dim() == 1 else y_true.shape[1]
    m = RegressionMetrics(cols, y_true.device, process_group)
    m.update(y_true, y_pred)
    return m.compute()


class ScoreSummary:
    """Streaming version of the result dict of evaluate_structural_prediction (unified_evaluator.py:393-405)."""

    def __init__(self, device, process_group=None):
        self.device, self.pg = torch.device(device), process_group
        self.sums = torch.zeros(6, device=self.device, dtype=torch.float64)
        self.ws = torch.empty(lib.pigan_eval_workspace_bytes(1), device=self.device, dtype=torch.uint8)
        self.n = 0

    def update(self, violations: Optional[torch.Tensor], recon_error: Optional[torch.Tensor],
               consistency: Optional[torch.Tensor]) -> None:
        v = None if violations is None else violations.to(torch.int32).contiguous()
        e = None if recon_error is None else _f32(recon_error, "recon_error")
        c = None if consistency is None else _f32(consistency, "consistency")
        n = next(x for x in (v, e, c) if x is not None).numel()
        check(lib.pigan_score_summary_sums(native.ptr(v), native.ptr(e), native.ptr(c), n, self.sums.data_ptr(), 1,
                                           self.ws.data_ptr(), self.ws.numel(), native.current_stream()))
        self.n += n

    def compute(self) -> Dict[str, float]:
        sums, n = self.sums, self.n
        if self.pg is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            sums = sums.clone()
            cnt = torch.tensor([n], device=self.device, dtype=torch.int64)
            dist.all_reduce(sums, group=self.pg)
            dist.all_reduce(cnt, group=self.pg)
            n = int(cnt.item())
        out = torch.empty(6, device=self.device, dtype=torch.float64)
        check(lib.pigan_score_summary_finalize(sums.data_ptr(), n, out.data_ptr(), native.current_stream()))
        res = dict(zip(SUMMARY_KEYS, out.cpu().tolist()))
        res["num_samples"] = n
        return res
