"""NVLink peer-memory exchange for data-parallel training (csrc/dp.cu, include/pigan_b200.h).

Every rank allocates one exchange region through the C ABI, the cudaIpc handles travel through
``torch.distributed.all_gather_object`` (plumbing), and the one-shot all-reduce kernels of the library sum the
batch-coupled buffers of the step straight out of the peers' memory.  ``DpExchange.create`` returns ``None`` when
peer mapping is not possible (single rank, different nodes, IPC refused) — the trainer then uses NCCL all-reduces.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import native
from .native import check, lib


class _DevMem:
    """torch view of raw device memory owned by the library (exported through __cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class DpExchange:
    def __init__(self, region: int, peers, ctx, rank: int, world: int, max_grad_floats: int, device):
        self.region, self.peers, self.ctx = region, peers, ctx
        self.rank, self.world, self.cap = rank, world, max_grad_floats
        self.device = device
        nbytes = lib.pigan_dp_region_bytes(max_grad_floats)
        self._bytes = torch.as_tensor(_DevMem(region, nbytes), device=device)

    @staticmethod
    def create(max_grad_floats: int, device, group=None) -> Optional["DpExchange"]:
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1 or world > 16 or os.environ.get("PIGAN_DP_EXCHANGE", "peer") != "peer":
            return None
        ok = True
        region = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        try:
            with torch.cuda.device(device):
                check(lib.pigan_dp_alloc(lib.pigan_dp_region_bytes(max_grad_floats), C.byref(region)))
                check(lib.pigan_dp_ipc_export(region, handle))
        except native.PiganError:
            ok = False
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, bytes(handle), os.uname().nodename), group=group)
        if not all(g[0] for g in gathered) or len({g[2] for g in gathered}) != 1:
            return None
        peers = (C.c_void_p * world)()
        try:
            with torch.cuda.device(device):
                for r in range(world):
                    if r == rank:
                        peers[r] = region.value
                    else:
                        p = C.c_void_p()
                        hb = (C.c_ubyte * 64).from_buffer_copy(gathered[r][1])
                        check(lib.pigan_dp_ipc_open(hb, C.byref(p)))
                        peers[r] = p.value
        except native.PiganError:
            ok = False
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            return None
        ctx = C.c_void_p()
        check(lib.pigan_dp_create(C.byref(ctx), world, rank, peers, max_grad_floats))
        return DpExchange(region.value, peers, ctx, rank, world, max_grad_floats, torch.device(device))

    def grad_slot(self, net: int, parity: int, n: int) -> torch.Tensor:
        """fp32 view [n] of gradient slot (net: 0 generator / 1 discriminator, parity) in this rank's region."""
        off = lib.pigan_dp_grad_slot_offset(self.cap, net, parity)
        return self._bytes[off:off + 4 * n].view(torch.float32)

    def allreduce_small(self, t: torch.Tensor, channel: int, epoch: int) -> None:
        is_double = 1 if t.dtype == torch.float64 else 0
        check(lib.pigan_dp_allreduce_small(self.ctx, t.data_ptr(), t.numel(), is_double, channel, epoch,
                                           native.current_stream()))

    def allreduce_small2(self, a: torch.Tensor, b: torch.Tensor, channel: int, epoch: int) -> None:
        """fp32 tensor a and fp64 tensor b in one exchange."""
        assert a.dtype == torch.float32 and b.dtype == torch.float64
        check(lib.pigan_dp_allreduce_small2(self.ctx, a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), channel, epoch,
                                            native.current_stream()))

    def allreduce_grads(self, net: int, dst: torch.Tensor, channel: int, epoch: int) -> None:
        check(lib.pigan_dp_allreduce_grads(self.ctx, net, dst.data_ptr(), dst.numel(), channel, epoch, None,
                                           native.current_stream()))
