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


class ForwardModel(nn.Module):
    def __init__(self, input_param_dim: int, output_spectrum_dim: int, output_metrics_dim: int):
        super().__init__()
        self.output_spectrum_dim = output_spectrum_dim
        self.output_metrics_dim = output_metrics_dim
        widths = (256, 512, 1024, 512, 256)
        layers = []
        fan_in = input_param_dim
        for w in widths:
            layers += [nn.Linear(fan_in, w), nn.LayerNorm(w), nn.LeakyReLU(0.2, inplace=True), nn.Dropout(0.2)]
            fan_in = w
        layers.append(nn.Linear(fan_in, output_spectrum_dim + output_metrics_dim))
        self.model = nn.Sequential(*layers)

    def forward(self, structural_params_norm: torch.Tensor):
        check_input(self, structural_params_norm, "structural_params_norm")
        if self.training:
            raise NotImplementedError("ForwardModel.forward in train() mode (active Dropout) is not on the native "
                                      "path; call .eval() as train_pigan does (core/train/train_pigan.py:75)")
        eng, flat = _pkg()
        st = flat.net_state(self, "forward_model")
        engine = eng.get_engine(structural_params_norm.device, structural_params_norm.shape[0])
        engine.load_forward_model(st.params.tensor())
        out = engine.forward_model_forward(structural_params_norm)
        return out[:, :self.output_spectrum_dim], out[:, self.output_spectrum_dim:]
