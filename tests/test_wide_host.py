"""Host-side checks of the widened dims (BASELINE config 5) that need no GPU: which PiganDims the library accepts,
the flat parameter count against the oracle's state_dict, and the drop-in module's constructor."""
import ctypes as C
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pi-gan-thz_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)


def test_accepted_dims():
    from pigan_b200 import native
    lib = native.lib
    ok = native.make_dims(spectrum_dim=2048, f_hidden=(2048,) * 5, g_hidden=(2048, 2048), d_hidden=(2048, 2048))
    assert lib.pigan_engine_workspace_bytes(C.byref(ok), 1024) > 0
    mixed = native.make_dims(spectrum_dim=500, f_hidden=(512, 1024, 2048, 1024, 256))
    assert lib.pigan_engine_workspace_bytes(C.byref(mixed), 1024) > 0
    assert lib.pigan_engine_workspace_bytes(C.byref(native.default_dims()), 1024) > 0
    for bad in (native.make_dims(f_hidden=(256, 512, 768, 512, 256)),      # width not 256/512/1024/2048
                native.make_dims(spectrum_dim=251),                         # odd spectrum length
                native.make_dims(spectrum_dim=4096, f_hidden=(2048,) * 5)):  # S + Mt beyond 2560
        assert lib.pigan_engine_workspace_bytes(C.byref(bad), 1024) == 0
    # a surrogate-only engine (spectrum length not a multiple of 64: no PI-GAN step) carries no generator /
    # discriminator activations; the config-5 engine carries the whole step's
    assert lib.pigan_engine_workspace_bytes(C.byref(mixed), 65536) < 3 << 30
    assert 4 << 30 < lib.pigan_engine_workspace_bytes(C.byref(ok), 65536) < 16 << 30


def test_param_count_matches_the_oracle_layout():
    from oracle import models as O
    from pigan_b200 import native
    hidden = (2048,) * 5
    dims = native.make_dims(spectrum_dim=2048, f_hidden=hidden)
    sd = O.init_forward_model(4, 2048, 8, hidden, gen=torch.Generator().manual_seed(0))
    n = sum(v.numel() for k, v in sd.items())
    assert native.lib.pigan_forward_model_param_count(C.byref(dims)) == n


def test_module_takes_widths():
    from core.models.forward_model import ForwardModel
    F = ForwardModel(4, 2048, 8, hidden=(2048,) * 5)
    assert F.model[0].out_features == 2048 and F.model[20].out_features == 2056
    assert list(F.state_dict()) == list(ForwardModel(4, 250, 8).state_dict())     # same keys as the reference stack
    assert ForwardModel(4, 250, 8).engine_dims() is None
    d = F.engine_dims()
    assert d.spectrum_dim == 2048 and tuple(d.f_hidden) == (2048,) * 5
    with pytest.raises(ValueError):
        ForwardModel(4, 250, 8, hidden=(256, 512))
