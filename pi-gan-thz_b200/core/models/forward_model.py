"""Drop-in ``core.models.forward_model.ForwardModel`` backed by the sm_100a engine.

Reference contract (core/models/forward_model.py:6-76): 4 normalised structure parameters -> (spectrum
[B,S], metrics [B,M]) returned as two views of one [B,S+M] buffer; ``state_dict`` keys
``model.{0,4,8,12,16,20}`` (Linear) and ``model.{1,5,9,13,17}`` (LayerNorm) so
forward_model_pretrained.pth loads unchanged.  The native forward is the eval-mode network (Dropout =
identity), which is how train_pigan and the evaluators use it; training the surrogate itself
(pretrain_fwd_model.py) is not on this path.
"""
import torch
import torch.nn as nn

from ._native import _pkg, check_input


class _SurrogateFn(torch.autograd.Function):
    """F(params_norm) with a backward pass into the INPUT (weights frozen): the vector-Jacobian product of the engine
    (pigan_forward_model_vjp).  This is how the reference's variant trainers use the surrogate - a loss on F(G(x))
    whose gradient reaches the generator through F (core/train/unified_trainer.py:240-256, 325)."""

    @staticmethod
    def forward(ctx, params_norm, engine, flat_params):
        engine.load_forward_model(flat_params)
        out = engine.forward_model_forward(params_norm)
        ctx.engine, ctx.flat = engine, flat_params
        ctx.save_for_backward(params_norm)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (p,) = ctx.saved_tensors
        dp = ctx.engine.forward_model_vjp(ctx.flat, p, grad_out.contiguous())
        ctx.engine.load_forward_model(ctx.flat)     # the VJP re-packs the weights: keep the frozen surrogate loaded
        return dp, None, None


class ForwardModel(nn.Module):
    REFERENCE_WIDTHS = (256, 512, 1024, 512, 256)

    def __init__(self, input_param_dim: int, output_spectrum_dim: int, output_metrics_dim: int, hidden=None):
        """``hidden``: five hidden widths (default: the reference's).  Other widths - 256/512/1024/2048 each, e.g.
        the widened surrogate of BASELINE config 5 with 2048-point spectra - run on an engine created for them."""
        super().__init__()
        self.output_spectrum_dim = output_spectrum_dim
        self.output_metrics_dim = output_metrics_dim
        widths = tuple(int(w) for w in (hidden if hidden is not None else self.REFERENCE_WIDTHS))
        if len(widths) != 5:
            raise ValueError("ForwardModel: five hidden widths")
        self.hidden = widths
        layers = []
        fan_in = input_param_dim
        for w in widths:
            layers += [nn.Linear(fan_in, w), nn.LayerNorm(w), nn.LeakyReLU(0.2, inplace=True), nn.Dropout(0.2)]
            fan_in = w
        layers.append(nn.Linear(fan_in, output_spectrum_dim + output_metrics_dim))
        self.model = nn.Sequential(*layers)

    def engine_dims(self):
        """None at the reference dims (the process-wide full engine), else the PiganDims of a surrogate-only engine."""
        if (self.hidden == self.REFERENCE_WIDTHS and self.output_spectrum_dim == 250 and self.output_metrics_dim == 8):
            return None
        from pigan_b200 import native
        return native.make_dims(spectrum_dim=self.output_spectrum_dim, metrics_dim=self.output_metrics_dim,
                                f_hidden=self.hidden)

    def forward(self, structural_params_norm: torch.Tensor):
        # under autograd the surrogate is differentiable with respect to its INPUT; its own parameters get no
        # gradient here (they are frozen wherever the reference back-propagates through F) - training F itself is
        # core.train.pretrain_fwd_model / pigan_fwd_train_step
        check_input(self, structural_params_norm, "structural_params_norm")
        if self.training:
            raise NotImplementedError("ForwardModel.forward in train() mode (active Dropout) is not on the native "
                                      "path; call .eval() as train_pigan does (core/train/train_pigan.py:75)")
        eng, flat = _pkg()
        st = flat.net_state(self, "forward_model")
        engine = eng.get_engine(structural_params_norm.device, structural_params_norm.shape[0], self.engine_dims())
        if torch.is_grad_enabled() and structural_params_norm.requires_grad:
            out = _SurrogateFn.apply(structural_params_norm.float().contiguous(), engine, st.params.tensor())
        else:
            engine.load_forward_model(st.params.tensor())
            out = engine.forward_model_forward(structural_params_norm)
        return out[:, :self.output_spectrum_dim], out[:, self.output_spectrum_dim:]
