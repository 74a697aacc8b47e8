"""Drop-in ``core.models.generator.Generator`` backed by the sm_100a engine.

Same constructor, ``state_dict`` keys (``main.{0,1,3,4,6}.*``) and forward contract as the reference
(core/models/generator.py:6-33): spectrum [B, input_dim] -> tanh-bounded normalised parameters
[B, output_dim]; BatchNorm uses batch statistics in ``train()`` mode (and updates its running buffers)
and running statistics in ``eval()`` mode.  The layer stack exists only to own the parameters (default
init consumes torch's RNG exactly like the reference); the math runs in csrc/ via the C ABI.
"""
import torch
import torch.nn as nn

from ._native import _pkg, check_input


class Generator(nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        widths = (512, 256)
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
        return engine.generator_forward(st.params.tensor(), st.bn.tensor(), st.nbt.tensor(), spectrum, self.training)
