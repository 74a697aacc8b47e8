"""Drop-in ``core.models.discriminator.Discriminator`` backed by the sm_100a engine.

Reference contract (core/models/discriminator.py:6-39): forward(spectrum [B,S], params [B,P]) ->
probability [B,1]; ``state_dict`` keys ``main.{0,2,4}.*`` with main.0.weight of shape [512, S+P]
(spectrum columns first).  The concatenation is never materialised: the parameter columns ride in spare
columns of the spectrum operand of the first tensor-core GEMM.
"""
import torch
import torch.nn as nn

from ._native import _pkg, check_input, module_params


class _DiscriminatorFn(torch.autograd.Function):
    """D(spectrum, params) with a backward pass into the module's parameters and into ``params`` (the path the
    generator's adversarial gradient takes in D(x, denormalize(G(x))), train_pigan.py:152-154):
    pigan_discriminator_backward recomputes the forward from the saved inputs, then runs the fused step's kernels."""

    @staticmethod
    def forward(ctx, spectrum, params_in, engine, st, *params):
        out = engine.discriminator_forward(st.params.tensor(), spectrum, params_in)
        ctx.engine, ctx.st = engine, st
        ctx.save_for_backward(spectrum, params_in)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, p = ctx.saved_tensors
        flat, gp = ctx.engine.discriminator_backward(ctx.st.params.tensor(), x, p, grad_out.contiguous(),
                                                     want_grad_params=ctx.needs_input_grad[1])
        return (None, gp, None, None, *ctx.st.params.views_like(flat))


class Discriminator(nn.Module):
    REFERENCE_WIDTHS = (512, 256)

    def __init__(self, input_spec_dim, input_param_dim, hidden=None):
        """``hidden``: the two hidden widths (default: the reference's 512, 256); see ``Generator``."""
        super().__init__()
        widths = tuple(int(w) for w in (hidden if hidden is not None else self.REFERENCE_WIDTHS))
        if len(widths) != 2:
            raise ValueError("Discriminator: two hidden widths")
        self.hidden = widths
        self.main = nn.Sequential(
            nn.Linear(input_spec_dim + input_param_dim, widths[0]), nn.LeakyReLU(0.2, inplace=True),
            nn.Linear(widths[0], widths[1]), nn.LeakyReLU(0.2, inplace=True),
            nn.Linear(widths[1], 1), nn.Sigmoid())

    def forward(self, spectrum, params):
        if spectrum.dim() > 2:
            spectrum = spectrum.view(spectrum.size(0), -1)
        if params.dim() > 2:
            params = params.view(params.size(0), -1)
        check_input(self, spectrum, "spectrum")
        check_input(self, params, "params")
        eng, flat = _pkg()
        st = flat.net_state(self, "discriminator")
        engine = eng.get_engine(spectrum.device, spectrum.shape[0])
        mp = module_params(st)
        if torch.is_grad_enabled() and (params.requires_grad or any(p.requires_grad for p in mp)):
            if spectrum.requires_grad:
                raise NotImplementedError("gradient with respect to the discriminator's input spectrum is not provided")
            return _DiscriminatorFn.apply(spectrum.float().contiguous(), params.float().contiguous(), engine, st, *mp)
        return engine.discriminator_forward(st.params.tensor(), spectrum, params)
