"""Shared plumbing of the three drop-in modules: route ``forward`` to the CUDA engine, fail loudly otherwise."""
from __future__ import annotations

import torch


def _pkg():
    import pigan_b200.engine as eng
    import pigan_b200.flat as flat
    return eng, flat


def check_input(module: torch.nn.Module, x: torch.Tensor, name: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{type(module).__name__}.forward: {name} is on {x.device}; the B200-native path has no "
                           "CPU fallback — move the module and its inputs to a CUDA device")
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()) and getattr(
            module, "_pigan_autograd_error", True):
        raise NotImplementedError(
            f"{type(module).__name__}.forward under autograd is not exposed by the native path: train with "
            "core.train.train_pigan.train_pigan (fused D-step/G-step kernels) or call the module under torch.no_grad()")
