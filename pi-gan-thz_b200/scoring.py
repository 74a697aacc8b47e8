"""Batched inverse-design scoring, sharded by candidate (BASELINE config 4).

The unit of work is one candidate: ``cand_i = target + sigma * noise_i`` (the reference's only candidate
sampler, core/evaluate/unified_evaluator.py:453-455), pushed through G(eval) -> F(eval) and scored with the
evaluator's per-sample reconstruction error ``mean((target - F(G(cand_i)).spectrum)**2)`` (:376-392).  Candidates
are cut into fixed-size chunks; chunk ``j`` always draws its noise from ``seed + j``, and ranks own contiguous
chunk ranges, so the ranking does not depend on the number of GPUs.  Every rank keeps a running top-k; the only
collective is one final all-gather of k (score, global index, 4 params) rows per rank followed by a local merge.

Two noise sources: ``noise="philox"`` (default) draws the noise inside the kernel from Philox4x32-10 keyed by the
seed and counted by the GLOBAL candidate index — one C call per rank (pigan_inverse_design_search), ranks own
contiguous index ranges, and candidate i gets the same noise at any world size.  ``noise="torch"`` feeds explicit
``torch.randn`` tensors chunk by chunk (the parity path).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import engine as _engine


def shard_chunks(num_chunks: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous chunk range [begin, end) of ``rank``; the first ``num_chunks % world`` ranks get one more."""
    base, extra = divmod(num_chunks, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous candidate range [begin, end) of ``rank`` (same rule as shard_chunks)."""
    return shard_chunks(n, rank, world)


def merge_topk(scores: torch.Tensor, indices: torch.Tensor, params: torch.Tensor, k: int,
               topk_fn=None):
    """k smallest of the concatenated partial results, ascending, ties by position (pigan_topk_smallest).
    ``topk_fn(scores, k) -> (values, positions)`` defaults to the CUDA kernel; CPU tests inject their own."""
    k = min(k, scores.numel())
    if topk_fn is None:
        pos_idx = torch.arange(scores.numel(), device=scores.device, dtype=torch.int64)
        vals, pos = _engine.topk_smallest(scores, k, in_indices=pos_idx)
    else:
        vals, pos = topk_fn(scores, k)
    return vals, indices[pos], params[pos]


def full_wave_chunk(device=None, waves: int = 4) -> int:
    """Candidates per chunk that tile the GPU without a partial wave: one 128-row tile per SM and wave (148 SMs ->
    75 776 for 4 waves).  A 65 536-candidate chunk is 3.46 waves, i.e. ~8 % of every GEMM's time runs on 68 of 148
    SMs; results do not depend on the chunk size (the noise is keyed by the global candidate index)."""
    sms = torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count
    return int(sms) * 128 * int(waves)


class InverseDesigner:
    def __init__(self, engine: _engine.Engine, g_flat: torch.Tensor, g_bn: torch.Tensor, chunk: Optional[int] = None,
                 process_group=None):
        self.engine = engine
        self.g_flat, self.g_bn = g_flat, g_bn
        self.chunk = int(chunk or engine.max_batch)
        if self.chunk > engine.max_batch:
            raise ValueError("chunk exceeds the engine's max_batch")
        self.pg = process_group
        self.topk_fn = None          # CPU tests inject a torch top-k; None = pigan_topk_smallest
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(process_group) if on else 0
        self.world = dist.get_world_size(process_group) if on else 1

    def search(self, target: torch.Tensor, num_candidates: int, k: int = 1024, sigma: float = 0.01,
               seed: int = 0, noise: str = "philox") -> Dict[str, torch.Tensor]:
        """Scores ``num_candidates`` noisy candidates around ``target`` [S] and returns the k best over all ranks
        (identical on every rank): ``recon_error`` [k] ascending, ``index`` [k] global candidate ids,
        ``params_norm`` [k,4].  ``scored`` is the number of candidates this rank processed."""
        dev = self.engine.device
        target = target.to(dev, torch.float32).reshape(-1).contiguous()
        S = target.numel()
        if noise == "philox":
            lo, hi = shard_range(num_candidates, self.rank, self.world)
            best_s, best_i, best_p = self.engine.search(self.g_flat, self.g_bn, target, sigma, seed, lo, hi - lo, k)
            keep = torch.isfinite(best_s)
            best_s, best_i, best_p = best_s[keep], best_i[keep], best_p[keep]
            if self.world > 1:
                best_s, best_i, best_p = self._gather_merge(best_s, best_i, best_p, k)
            return {"recon_error": best_s, "index": best_i, "params_norm": best_p, "scored": hi - lo}
        if noise != "torch":
            raise ValueError("noise must be 'philox' or 'torch'")
        num_chunks = (num_candidates + self.chunk - 1) // self.chunk
        c0, c1 = shard_chunks(num_chunks, self.rank, self.world)
        best_s = torch.empty(0, device=dev, dtype=torch.float32)
        best_i = torch.empty(0, device=dev, dtype=torch.int64)
        best_p = torch.empty(0, 4, device=dev, dtype=torch.float32)
        gen = torch.Generator(device=dev)
        noise = torch.empty(self.chunk, S, device=dev, dtype=torch.float32)
        scored = 0
        for j in range(c0, c1):
            base = j * self.chunk
            n = min(self.chunk, num_candidates - base)
            gen.manual_seed(seed + j)
            noise.normal_(generator=gen)
            out = self.engine.score_candidates(self.g_flat, self.g_bn, target=target, noise=noise[:n], sigma=sigma)
            kk = min(k, n)
            s, i = _engine.topk_smallest(out["recon_error"], kk, index_base=base)
            p = out["params_norm"][i - base]
            best_s, best_i, best_p = merge_topk(torch.cat([best_s, s]), torch.cat([best_i, i]),
                                                torch.cat([best_p, p]), k)
            scored += n
        if self.world > 1:
            best_s, best_i, best_p = self._gather_merge(best_s, best_i, best_p, k)
        return {"recon_error": best_s, "index": best_i, "params_norm": best_p, "scored": scored}

    def _gather_merge(self, s, i, p, k):
        """The one collective of the scoring path: all-gather of each rank's k best rows, then a local merge."""
        dev = s.device
        row = torch.full((k, 6), float("inf"), device=dev, dtype=torch.float64)
        n = s.numel()
        row[:n, 0] = s.double()
        row[:n, 1] = i.double()          # exact below 2^53 candidates
        row[:n, 2:6] = p.double()
        allrows = torch.empty(self.world * k, 6, device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allrows, row, group=self.pg)
        valid = torch.isfinite(allrows[:, 0])
        allrows = allrows[valid]
        return merge_topk(allrows[:, 0].float().contiguous(), allrows[:, 1].long().contiguous(),
                          allrows[:, 2:6].float().contiguous(), k, topk_fn=self.topk_fn)
