"""Host side of the forward-surrogate training step: mirrors the loop body of the reference's
``pretrain_forward_model`` (core/train/pretrain_fwd_model.py:68-92) on top of ``pigan_fwd_train_step``.

The module's parameters are re-pointed at one flat fp32 buffer in ``state_dict`` order (flat.py), Adam's moments and
the gradients live in buffers of the same layout, and one C call per batch runs forward (train mode, counter-based
Dropout), the two MSE losses, backward, clip_grad_norm_(1.0) and Adam.  Data parallel = replicas: every rank runs
phase 0 on its rows, gradients and loss sums are all-reduced (NCCL), phase 1 applies the identical update.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import engine as _engine
from . import flat as _flat
from . import native
from .native import PiganFwdTrainArgs, check, lib

LOSS_KEYS = ("loss", "loss_spectrum", "loss_metrics")


class ForwardTrainer:
    def __init__(self, forward_model, device, max_batch: int, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_norm: float = 1.0, dropout_p: float = 0.2, seed: int = 0, process_group=None,
                 engine: Optional[_engine.Engine] = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ForwardTrainer needs a CUDA device — the B200 path has no CPU fallback")
        self.f = forward_model.to(self.device)
        self.fs = _flat.net_state(forward_model, "forward_model")
        if engine is None:
            dims = forward_model.engine_dims() if hasattr(forward_model, "engine_dims") else None
            engine = _engine.Engine(max_batch, self.device, dims)   # widened dims: a surrogate-only engine
        self.engine = engine
        fp = self.fs.params.tensor()
        self.grads, self.m, self.v = (torch.zeros_like(fp) for _ in range(3))
        self.losses = torch.zeros(3, device=self.device, dtype=torch.float32)
        self.loss_sums = torch.zeros(2, device=self.device, dtype=torch.float32)
        self.workspace = torch.empty(lib.pigan_fwd_train_workspace_bytes(self.engine.handle), dtype=torch.uint8,
                                     device=self.device)
        self.betas, self.eps, self.max_norm = betas, float(eps), float(max_norm)
        self.dropout_p, self.seed = float(dropout_p), int(seed)
        self.step_count = 0
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        for p, gview in zip(self.fs.params._tensors(), self.fs.params.views_like(self.grads)):
            p.grad = gview   # where autograd would have left them (clipped, as after clip_grad_norm_)

    def hidden_total(self) -> int:
        return int(sum(self.engine.dims.f_hidden))

    def step(self, params_norm: torch.Tensor, spectrum: torch.Tensor, metrics_norm: torch.Tensor, lr: float,
             first_row: Optional[int] = None, mask_dump: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimiser step on device-resident fp32 tensors; returns the [3] device tensor (total, spectrum,
        metrics loss — means over the global batch).  ``first_row``: global index of this rank's first row, the
        Dropout counter (default rank * batch).  ``mask_dump``: uint8 [sum(H_i) * batch] receiving the keep-masks."""
        for name, x in (("params_norm", params_norm), ("spectrum", spectrum), ("metrics_norm", metrics_norm)):
            if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
                raise RuntimeError(f"ForwardTrainer.step: {name} must be a contiguous fp32 CUDA tensor")
        B = params_norm.shape[0]
        if self.world > 1:
            from .trainer import check_equal_batch
            self._dp_batch = check_equal_batch(B, getattr(self, "_dp_batch", None), self.pg)
        self.step_count += 1
        a = PiganFwdTrainArgs()
        a.params_norm, a.spectrum, a.metrics_norm = params_norm.data_ptr(), spectrum.data_ptr(), metrics_norm.data_ptr()
        a.batch, a.global_batch = B, B * self.world
        a.first_row = self.rank * B if first_row is None else int(first_row)
        a.f_params = self.fs.params.tensor().data_ptr()
        a.f_grads, a.f_exp_avg, a.f_exp_avg_sq = self.grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr()
        a.lr, a.step = float(lr), self.step_count
        a.beta1, a.beta2, a.eps, a.max_norm = float(self.betas[0]), float(self.betas[1]), self.eps, self.max_norm
        a.dropout_p, a.dropout_seed = self.dropout_p, self.seed
        a.losses, a.loss_sums = self.losses.data_ptr(), self.loss_sums.data_ptr()
        a.mask_dump = native.ptr(mask_dump)
        ws, nb, st = self.workspace.data_ptr(), self.workspace.numel(), native.current_stream()
        if self.world == 1:
            check(lib.pigan_fwd_train_step(self.engine.handle, C.byref(a), ws, nb, st))
        else:
            check(lib.pigan_fwd_train_step_phase(self.engine.handle, C.byref(a), 0, ws, nb, st))
            dist.all_reduce(self.grads, group=self.pg)
            dist.all_reduce(self.loss_sums, group=self.pg)
            check(lib.pigan_fwd_train_step_phase(self.engine.handle, C.byref(a), 1, ws, nb, st))
        return self.losses
