"""Drop-in ``core.models.generator.Generator`` backed by the sm_100a engine.

Same constructor, ``state_dict`` keys (``main.{0,1,3,4,6}.*``) and forward contract as the reference
(core/models/generator.py:6-33): spectrum [B, input_dim] -> tanh-bounded normalised parameters
[B, output_dim]; BatchNorm uses batch statistics in ``train()`` mode (and updates its running buffers)
and running statistics in ``eval()`` mode.  The layer stack exists only to own the parameters (default
init consumes torch's RNG exactly like the reference); the math runs in csrc/ via the C ABI.
"""
import torch
import torch.nn as nn

from ._native import _pkg, check_input, module_params


class _GeneratorFn(torch.autograd.Function):
    """G(spectrum) in train mode with a backward pass into the module's parameters (pigan_generator_backward: the
    forward is recomputed from the saved spectrum, then the fused step's backward kernels run).  No gradient with
    respect to the spectrum (the reference never asks for one)."""

    @staticmethod
    def forward(ctx, spectrum, engine, st, *params):
        out = engine.generator_forward(st.params.tensor(), st.bn.tensor(), st.nbt.tensor(), spectrum, True)
        ctx.engine, ctx.st = engine, st
        ctx.save_for_backward(spectrum)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        flat = ctx.engine.generator_backward(ctx.st.params.tensor(), x, grad_out.contiguous())
        return (None, None, None, *ctx.st.params.views_like(flat))


class Generator(nn.Module):
    REFERENCE_WIDTHS = (512, 256)

    def __init__(self, input_dim, output_dim, hidden=None):
        """``hidden``: the two hidden widths (default: the reference's 512, 256).  Other widths (BASELINE config 5:
        2048, 2048 on 2048-point spectra) train through ``NativeTrainer`` - the fused step - only; the stand-alone
        module forward / backward entry points exist at the reference widths."""
        super().__init__()
        widths = tuple(int(w) for w in (hidden if hidden is not None else self.REFERENCE_WIDTHS))
        if len(widths) != 2:
            raise ValueError("Generator: two hidden widths")
        self.hidden = widths
        self.main = nn.Sequential(
            nn.Linear(input_dim, widths[0]), nn.BatchNorm1d(widths[0]), nn.ReLU(True),
            nn.Linear(widths[0], widths[1]), nn.BatchNorm1d(widths[1]), nn.ReLU(True),
            nn.Linear(widths[1], output_dim), nn.Tanh())

    def forward(self, spectrum):
        if spectrum.dim() > 2:
            spectrum = spectrum.view(spectrum.size(0), -1)
        check_input(self, spectrum, "spectrum")
        eng, flat = _pkg()
        st = flat.net_state(self, "generator")
        engine = eng.get_engine(spectrum.device, spectrum.shape[0])
        params = module_params(st)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            if not self.training:
                raise NotImplementedError("Generator.forward under autograd in eval() mode (BatchNorm on running "
                                          "statistics) is not on the native path; the reference differentiates the "
                                          "generator in train() mode only (core/train/train_pigan.py:117)")
            if spectrum.requires_grad:
                raise NotImplementedError("gradient with respect to the generator's input spectrum is not provided")
            return _GeneratorFn.apply(spectrum.float().contiguous(), engine, st, *params)
        return engine.generator_forward(st.params.tensor(), st.bn.tensor(), st.nbt.tensor(), spectrum, self.training)
