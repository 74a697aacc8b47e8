"""Drop-in ``core.utils.loss`` — the six loss factories/functions of the reference (core/utils/loss.py:8-147)
with identical signatures, for callers that compose losses themselves.  Inside ``train_pigan`` these are not
called: the fused CUDA step evaluates the same formulas in the GEMM epilogues (csrc/epilogues.cuh)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def criterion_bce():
    return nn.BCELoss()          # D ends in a Sigmoid (loss.py:17)


def criterion_mse():
    return nn.MSELoss()          # loss.py:25


def maxwell_equation_loss(predicted_spectrum, frequencies, predicted_params_norm):
    """Mean squared second difference along the frequency axis (loss.py:45-64); the other two arguments are
    accepted and unused, as in the reference."""
    if predicted_spectrum.size(1) < 3:
        return torch.zeros(1, device=predicted_spectrum.device)
    s = predicted_spectrum
    first = s[:, 1:] - s[:, :-1]
    second = first[:, 1:] - first[:, :-1]   # difference of differences, the reference's evaluation order
    return torch.mean(second ** 2)


def lc_model_approx_loss(f1_pred_norm, f2_pred_norm, structural_params_norm):
    """MSE of the predicted (normalised) resonances against 0.4 r1 + 0.6 w and 0.3 r2 + 0.7 g (loss.py:82-101)."""
    p = structural_params_norm
    target1 = 0.4 * p[:, 0:1] + 0.6 * p[:, 2:3]
    target2 = 0.3 * p[:, 1:2] + 0.7 * p[:, 3:4]
    return F.mse_loss(f1_pred_norm, target1) + F.mse_loss(f2_pred_norm, target2)


def structural_param_range_loss(predicted_params_norm):
    """Quadratic penalty outside [0, 1] (loss.py:121-127)."""
    below = torch.clamp(0 - predicted_params_norm, min=0) ** 2
    above = torch.clamp(predicted_params_norm - 1, min=0) ** 2
    return torch.mean(below + above)


def bnn_kl_loss(model: nn.Module):
    """Constant zero of shape (1,) (loss.py:147) — which is why loss_g_total has shape (1,)."""
    return torch.zeros(1, device=next(model.parameters()).device)
