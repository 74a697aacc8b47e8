"""Shared plumbing of the three drop-in modules: route ``forward`` to the CUDA engine, fail loudly otherwise.

Under autograd the modules behave like the reference's: ``Generator`` and ``Discriminator`` return tensors with a
``grad_fn`` (``torch.autograd.Function`` wrappers over pigan_generator_backward / pigan_discriminator_backward) whose
backward fills the ``.grad`` of the module's parameters - and of the discriminator's ``params`` input, which is how
the generator's adversarial gradient flows in ``D(x, denormalize(G(x)))``; ``ForwardModel`` is differentiable in its
input with frozen weights (pigan_forward_model_vjp).  The fused ``train_pigan`` step does not go through these."""
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


def module_params(st):
    """The module's nn.Parameters in flat-buffer order (passed to the autograd Functions so that autograd routes
    their gradients)."""
    return st.params._tensors()
