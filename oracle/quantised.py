"""ORACLE (test infrastructure): quantisation-aware fp64 restatement of the PI-GAN train step's gradients.

Purpose: separate ARITHMETIC error of the B200 path from the CONDITIONING of the gradients it is compared on.
``train_step_grads(..., quantise=True)`` evaluates train_pigan.py:123-187 in float64 but rounds the FORWARD values to
fp16 exactly where the CUDA engine stores or feeds fp16 operands (pi-gan-thz_b200/csrc/engine.cu): the centred
spectrum operand, every weight copy that enters a tensor-core GEMM, and the activations between the layers.  The
backward pass is exact float64 autograd of that rounded forward pass (straight-through rounding, which is what the
engine's backward kernels compute: they read the stored, rounded activations).  ``quantise=False`` is the plain
float64 step.  The relative distance between the two is the deviation ANY implementation with fp16 operands and
otherwise infinite precision shows against the reference - the floor the GPU parity tests must be read against
(tests/test_quantisation_floor.py pins it; tests/test_gpu_engine.py derives its gradient tolerances from it).

Never imported by the product.
"""
from __future__ import annotations

import copy
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import models as O

Tensor = torch.Tensor


SITES = ("input", "weights", "activations")   # what is rounded when quantise=True (ablations narrow this down)
_enabled = set(SITES)


def set_sites(sites) -> None:
    """Restrict the rounding to a subset of SITES (ablation: which rounding costs what)."""
    _enabled.clear()
    _enabled.update(sites)


def q16(x: Tensor, on: bool, site: str = "activations") -> Tensor:
    """Round to fp16 in the forward pass, identity in the backward pass."""
    if not on or site not in _enabled:
        return x
    return x + (x.detach().to(torch.float16).to(x.dtype) - x.detach())


def _bn_train(h: Tensor, gamma: Tensor, beta: Tensor) -> Tensor:
    mu = h.mean(0)
    var = h.var(0, unbiased=False)
    return (h - mu) / torch.sqrt(var + O.BN_EPS) * gamma + beta


def generator_forward(g: Dict[str, Tensor], xq: Tensor, on: bool) -> Tensor:
    """generator.py:28-33 in training mode on the centred operand xq (BatchNorm is invariant to the constant shift
    the centring and the first bias add; the engine stores the pre-BatchNorm products without them)."""
    h1 = q16(xq @ q16(g["main.0.weight"], on, "weights").t(), on)
    a1 = q16(torch.relu(_bn_train(h1, g["main.1.weight"], g["main.1.bias"])), on)
    h2 = q16(a1 @ q16(g["main.3.weight"], on, "weights").t(), on)
    a2 = torch.relu(_bn_train(h2, g["main.4.weight"], g["main.4.bias"]))
    return torch.tanh(a2 @ g["main.6.weight"].t() + g["main.6.bias"])


def discriminator_forward(d: Dict[str, Tensor], xq: Tensor, c: Tensor, params: Tensor, on: bool) -> Tensor:
    """discriminator.py:30-39 on the engine's operand layout: [xq | params - 2.5 | 1], the centring row c and the
    2.5 folded into an effective bias in full precision."""
    S = xq.shape[1]
    w1 = d["main.0.weight"]
    b_eff = d["main.0.bias"] + w1[:, :S] @ c + 2.5 * w1[:, S:].sum(1)
    z1 = xq @ q16(w1[:, :S], on, "weights").t() + q16(params - 2.5, on, "input") @ q16(w1[:, S:], on, "weights").t() + b_eff
    z1 = q16(F.leaky_relu(z1, O.LEAKY), on)
    z2 = F.leaky_relu(z1 @ q16(d["main.2.weight"], on, "weights").t() + d["main.2.bias"], O.LEAKY)
    return torch.sigmoid(z2 @ d["main.4.weight"].t() + d["main.4.bias"])


def forward_model_forward(f: Dict[str, Tensor], p: Tensor, S: int, on: bool, stored_pre_ln: bool = False
                          ) -> Tuple[Tensor, Tensor]:
    """forward_model.py:62-76 (eval): fp16 activations between the layers, fp16 weight copies from layer 2 on.
    stored_pre_ln: the backward-capable path of the engine (pigan_fwd_train_step, pigan_forward_model_input_grad /
    _vjp) stores the Linear output of the tensor-core layers in fp16 and normalises the stored values; the inference
    path (EpiLnStore) normalises the fp32 accumulator."""
    h = p
    for k, (li, ni) in enumerate(zip(O.F_LINEAR[:-1], O.F_NORM)):
        w = f[f"model.{li}.weight"]
        h = h @ (w if k == 0 else q16(w, on, "weights")).t() + f[f"model.{li}.bias"]
        if stored_pre_ln and k > 0:
            h = q16(h, on)
        h = F.layer_norm(h, (h.shape[1],), f[f"model.{ni}.weight"], f[f"model.{ni}.bias"], O.LN_EPS)
        h = q16(F.leaky_relu(h, O.LEAKY), on)
    out = h @ q16(f["model.20.weight"], on, "weights").t() + f["model.20.bias"]
    return out[:, :S], out[:, S:]


def train_step_grads(g_sd, d_sd, f_sd, batch, quantise: bool, lr_d: float = 2e-4, f1_idx: int = 0, f2_idx: int = 1):
    """Unclipped float64 gradients of the D-step and of the G-step (against the updated discriminator) of one
    train_pigan iteration.  Returns (d_grads, g_grads) keyed like the state dicts."""
    dt = torch.float64
    g_sd, d_sd, f_sd = (O.cast_state(copy.deepcopy(s), dt) for s in (g_sd, d_sd, f_sd))
    spec, praw, _pn, _md, mnorm = batch
    spec, praw, mnorm = spec.to(dt), praw.to(dt), mnorm.to(dt)
    B, S = spec.shape
    on = quantise
    c = spec[: min(B, 512)].mean(0)          # engine: launch_center_vec over the first <= 512 rows
    if on:
        c = c.float().double()               # the centring row is an fp32 vector
    xq = q16(spec - c, on, "input")

    # ---- D-step (train_pigan.py:123-143)
    d = O._leaf(d_sd, O.D_TRAINABLE)
    out_real = discriminator_forward(d, xq, c, praw, on)
    with torch.no_grad():
        pden = O.denormalize_params(generator_forward(g_sd, xq, on))
    out_fake = discriminator_forward(d, xq, c, pden, on)
    loss_d = O.bce(out_real, torch.full((B, 1), 0.9, dtype=dt)) + O.bce(out_fake, torch.full((B, 1), 0.1, dtype=dt))
    dl = torch.autograd.grad(loss_d, [d[n] for n in O.D_TRAINABLE])
    d_grads = {n: x.clone() for n, x in zip(O.D_TRAINABLE, dl)}
    clipped = {n: x.clone() for n, x in d_grads.items()}
    O.clip_grad_norm_(list(clipped.values()), 1.0)
    with torch.no_grad():
        O.Adam(O.D_TRAINABLE).step(d_sd, clipped, lr_d)

    # ---- G-step (train_pigan.py:145-187)
    g = O._leaf(g_sd, O.G_TRAINABLE)
    p = generator_forward(g, xq, on)
    out_g = discriminator_forward(d_sd, xq, c, O.denormalize_params(p), on)
    loss_adv = O.bce(out_g, torch.ones(B, 1, dtype=dt))
    with torch.no_grad():
        recon, pm = forward_model_forward(f_sd, p, S, on)
    loss_lc = O.lc_model_approx_loss(pm[:, f1_idx:f1_idx + 1], pm[:, f2_idx:f2_idx + 1], p)
    loss_range = O.structural_param_range_loss(p)
    loss_g = loss_adv + O.LAMBDA_LC * loss_lc + O.LAMBDA_PARAM_RANGE * loss_range   # the terms that carry gradient
    # main.0.bias / main.3.bias feed a train-mode BatchNorm: their gradient is exactly zero (not part of the graph here)
    gl = torch.autograd.grad(loss_g, [g[n] for n in O.G_TRAINABLE], allow_unused=True)
    g_grads = {n: (torch.zeros_like(g[n]) if x is None else x.clone()) for n, x in zip(O.G_TRAINABLE, gl)}
    return d_grads, g_grads


def relative_floor(g_sd, d_sd, f_sd, batch) -> Dict[str, Dict[str, float]]:
    """Per-tensor relative L2 distance between the fp16-forward and the exact float64 gradients."""
    de, ge = train_step_grads(g_sd, d_sd, f_sd, batch, quantise=False)
    dq, gq = train_step_grads(g_sd, d_sd, f_sd, batch, quantise=True)

    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-300))

    return {"d": {n: rel(dq[n], de[n]) for n in de}, "g": {n: rel(gq[n], ge[n]) for n in ge}}
