"""TEST INFRASTRUCTURE — CPU restatement (numpy, float64) of the evaluator's reductions.  Pinned to the reference by
tests/golden/evaluator_metrics.npz (tools/make_golden.py runs UnifiedEvaluator.calculate_metrics itself, i.e.
sklearn + scipy, and numpy's mean/std).  Only tests/ may import this.

  regression_metrics : core/evaluate/unified_evaluator.py:138-184
  score_summary      : core/evaluate/unified_evaluator.py:393-405
"""
import numpy as np


def regression_metrics(y_true, y_pred):
    y32, p32 = np.asarray(y_true, dtype=np.float32), np.asarray(y_pred, dtype=np.float32)
    if y32.ndim == 1:
        y32, p32 = y32[:, None], p32[:, None]
    y, p = y32.astype(np.float64), p32.astype(np.float64)
    d = y - p
    mse = float(np.mean(d * d))                       # :152  mean_squared_error (uniform average over columns)
    mae = float(np.mean(np.abs(d)))                   # :153
    ss_res = np.sum(d * d, axis=0)
    ss_tot = np.sum((y - y.mean(axis=0)) ** 2, axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        r2c = np.where(ss_tot > 0, 1.0 - ss_res / ss_tot, np.where(ss_res == 0, 1.0, 0.0))   # :158 r2_score, force_finite
        yc, pc = y - y.mean(axis=0), p - p.mean(axis=0)
        den = np.sqrt(np.sum(yc * yc, axis=0) * np.sum(pc * pc, axis=0))
        prc = np.where(den > 0, np.sum(yc * pc, axis=0) / den, np.nan)                        # :165-176 pearsonr per column
    ape = np.abs((y32 - p32) / (y32 + np.float32(1e-8)))                                      # :182, float32 like numpy
    return {"mse": mse, "mae": mae, "rmse": float(np.sqrt(mse)), "r2": float(np.mean(r2c)),
            "pearson_r": float(np.mean(prc)), "mape": float(np.mean(ape.astype(np.float64)) * 100.0)}


def score_summary(violations, recon_error, consistency):
    v = np.asarray(violations)
    e, c = np.asarray(recon_error, dtype=np.float64), np.asarray(consistency, dtype=np.float64)
    return {"param_range_violation_rate": float(np.mean(v > 0)), "avg_param_violations": float(np.mean(v)),
            "reconstruction_error_mean": float(np.mean(e)), "reconstruction_error_std": float(np.std(e)),
            "consistency_score_mean": float(np.mean(c)), "consistency_score_std": float(np.std(c)),
            "num_samples": int(v.shape[0])}
