"""Multi-process host logic on CPU (gloo, world_size 2): the data-parallel phase schedule of the trainer and the
one-collective top-k merge of sharded scoring.  The CUDA engine is replaced by recording stand-ins — the kernels
themselves are covered by the -m gpu tests (incl. a two-rank emulation on one GPU that uses the same schedule)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pi-gan-thz_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _schedule_worker(rank, world, port, out):
    _init(rank, world, port)
    from pigan_b200.trainer import dp_phase_plan, run_dp_step
    h1, h2 = 512, 256
    g = torch.Generator().manual_seed(100 + rank)
    bufs = {"bn_sums": torch.rand(2 * h1 + 2 * h2, generator=g), "bn_bwd_sums": torch.rand(2 * h1 + 2 * h2, generator=g),
            "d_grads": torch.rand(1000, generator=g), "g_grads": torch.rand(1200, generator=g),
            "loss_sums": torch.rand(16, generator=g, dtype=torch.float64)}
    before = {k: v.clone() for k, v in bufs.items()}
    log = []

    def run_phase(ph):
        # a phase may only rely on reductions of EARLIER phases: record what has been reduced so far
        log.append(ph)

    run_dp_step(run_phase, lambda name: bufs[name], lambda t: dist.all_reduce(t), dp_phase_plan(h1, h2))
    gathered = {}
    for k, v in before.items():
        parts = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(parts, v)
        gathered[k] = sum(parts)
    ok = log == list(range(7))
    for k in ("bn_sums", "bn_bwd_sums", "d_grads", "g_grads"):
        ok = ok and torch.allclose(bufs[k], gathered[k])
    ok = ok and torch.allclose(bufs["loss_sums"][:8], gathered["loss_sums"][:8])
    ok = ok and torch.equal(bufs["loss_sums"][8:], before["loss_sums"][8:])      # grad-norm slots stay local
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _merge_worker(rank, world, port, out):
    _init(rank, world, port)
    from pigan_b200 import scoring
    k, per_rank = 16, 1000
    g = torch.Generator().manual_seed(7)
    all_scores = torch.rand(world * per_rank, generator=g)
    all_params = torch.rand(world * per_rank, 4, generator=g)
    lo, hi = rank * per_rank, (rank + 1) * per_rank

    def cpu_topk(s, kk):
        order = torch.argsort(s.double(), stable=True)[:kk]
        return s[order], order

    des = scoring.InverseDesigner.__new__(scoring.InverseDesigner)
    des.pg, des.rank, des.world = None, rank, world
    des.topk_fn = cpu_topk
    s, pos = cpu_topk(all_scores[lo:hi], k)
    bs, bi, bp = des._gather_merge(s, pos + lo, all_params[lo:hi][pos], k)
    es, epos = cpu_topk(all_scores, k)
    out[rank] = bool(torch.equal(bs, es) and torch.equal(bi, epos) and torch.equal(bp, all_params[epos]))
    dist.destroy_process_group()


def _spawn(fn, port):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(fn, args=(world, port, out), nprocs=world, join=True)
    return dict(out)


def test_dp_schedule_reduces_every_buffer_once_in_phase_order():
    assert _spawn(_schedule_worker, 29611) == {0: True, 1: True}


def test_sharded_topk_gather_merge_matches_global_topk():
    assert _spawn(_merge_worker, 29612) == {0: True, 1: True}


def test_shard_chunks_partition():
    from pigan_b200 import scoring
    for chunks in (0, 1, 7, 8, 1526):
        for world in (1, 2, 4, 8):
            spans = [scoring.shard_chunks(chunks, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == chunks
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
