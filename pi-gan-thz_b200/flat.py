"""Flat fp32 parameter buffers behind nn.Module parameters.

The C ABI takes each network's trainable tensors as ONE contiguous device buffer in state_dict order
(include/pigan_b200.h).  ``FlatParams`` builds that buffer from a module and re-points every
``nn.Parameter`` / buffer at a view of it, so ``state_dict()``, ``load_state_dict()``, checkpoints and
optimiser state keep working on the very memory the kernels read and write.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


class FlatParams:
    def __init__(self, module: torch.nn.Module, names: Sequence[str]):
        self.module = module
        self.names: List[str] = list(names)
        self.flat: torch.Tensor | None = None
        self.offsets: List[int] = []
        self._sync()

    def _tensors(self):
        lookup = dict(self.module.named_parameters())
        lookup.update(dict(self.module.named_buffers()))
        return [lookup[n] for n in self.names]

    def _is_synced(self, tensors) -> bool:
        if self.flat is None:
            return False
        base = self.flat.data_ptr()
        es = self.flat.element_size()
        for t, off in zip(tensors, self.offsets):
            if t.device != self.flat.device or t.data_ptr() != base + off * es or not t.is_contiguous():
                return False
        return True

    def _sync(self) -> None:
        tensors = self._tensors()
        if self._is_synced(tensors):
            return
        dev = tensors[0].device
        dtype = tensors[0].dtype
        self.offsets = []
        total = 0
        for t in tensors:
            self.offsets.append(total)
            total += t.numel()
        flat = torch.empty(total, dtype=dtype, device=dev)
        with torch.no_grad():
            for t, off in zip(tensors, self.offsets):
                flat[off:off + t.numel()].copy_(t.detach().reshape(-1))
                t.data = flat[off:off + t.numel()].view(t.shape)
        self.flat = flat

    def tensor(self) -> torch.Tensor:
        """The flat buffer (rebuilt if the module was moved / re-assigned since the last call)."""
        self._sync()
        return self.flat

    def views_like(self, other: torch.Tensor):
        """Per-parameter views of another flat buffer with the same layout (grads, Adam moments)."""
        tensors = self._tensors()
        return [other[off:off + t.numel()].view(t.shape) for t, off in zip(tensors, self.offsets)]


GENERATOR_PARAMS = ["main.0.weight", "main.0.bias", "main.1.weight", "main.1.bias", "main.3.weight", "main.3.bias",
                    "main.4.weight", "main.4.bias", "main.6.weight", "main.6.bias"]
GENERATOR_BN = ["main.1.running_mean", "main.1.running_var", "main.4.running_mean", "main.4.running_var"]
GENERATOR_NBT = ["main.1.num_batches_tracked", "main.4.num_batches_tracked"]
DISCRIMINATOR_PARAMS = ["main.0.weight", "main.0.bias", "main.2.weight", "main.2.bias", "main.4.weight", "main.4.bias"]
FORWARD_PARAMS = [f"model.{i}.{s}" for i in (0, 1, 4, 5, 8, 9, 12, 13, 16, 17, 20) for s in ("weight", "bias")]


class NetState:
    """Flat parameter (+ BatchNorm buffer) views of one of the three reference networks."""

    def __init__(self, module: torch.nn.Module, kind: str):
        self.kind = kind
        self.module = module
        if kind == "generator":
            self.params = FlatParams(module, GENERATOR_PARAMS)
            self.bn = FlatParams(module, GENERATOR_BN)
            self.nbt = FlatParams(module, GENERATOR_NBT)
        elif kind == "discriminator":
            self.params = FlatParams(module, DISCRIMINATOR_PARAMS)
            self.bn = self.nbt = None
        elif kind == "forward_model":
            self.params = FlatParams(module, FORWARD_PARAMS)
            self.bn = self.nbt = None
        else:
            raise ValueError(kind)


def net_state(module: torch.nn.Module, kind: str) -> NetState:
    st = getattr(module, "_pigan_state", None)
    if st is None or st.kind != kind:
        st = NetState(module, kind)
        object.__setattr__(module, "_pigan_state", st)
    return st
