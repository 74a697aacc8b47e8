#!/usr/bin/env python
"""Benchmark of the PI-GAN-THz hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch B] [--config wide]

Native arm: one "step" is one D-step + G-step of train_pigan's inner loop (core/train/train_pigan.py:114-187) at
batch B per GPU (default 65 536 = BASELINE config 2) on synthetic spectra of the dataset's shape, through the C ABI
(libpigan_b200.so).  `value` = samples/s with the batch resident in HBM; `e2e` = the same step fed from pinned HOST
buffers (H2D of the batch and D2H of the 9 losses inside the timed region).  The JSON line also carries the
inverse-design scoring throughput (candidates/s, `scoring`), the roofline of the dominant kernel measured with CUDA
events inside the timed region, a CPU baseline (the oracle port of the reference step on the host cores) and the
clocks seen during the timed region.

--config wide: the line is the same step at the BASELINE config-5 widths (generator 2048-2048-2048-4, discriminator
2052-2048-2048-1, surrogate 4-2048x5-2056, 2048-point spectra; 193 MFLOP per sample) with its own roofline, e2e and
CPU-port legs, and the widened surrogate-training step beside it; the default line carries both as side blocks
(`wide_pigan_training`, `wide_surrogate_training`).

Reference arm (--impl reference): the reference's CPU PyTorch path — restated in oracle/models.py and pinned to the
reference itself by tests/golden — timed on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))

import torch

# algorithmic work per unit: minimal necessary FLOPs
FLOP_PER_TRAIN_SAMPLE = 7_465_984
FLOP_PER_CANDIDATE = 3_275_776
FLOP_PER_PRETRAIN_SAMPLE = 2 * 4_132_352   # F fwd + bwd (dW + dX), MACs x 2
METRIC = "PI-GAN train samples/s"
# DRAM traffic of the dominant kernel per launch from the committed ncu capture (profiles/r02_gemm_kernels_ncu_full.csv,
# round 2, the kernels as they are now): the four Linear+LayerNorm launches of the forward surrogate read+write
# 50.8 + 150.7 + 174.6 + 74.7 MB (algorithmic fp16 activation bytes: 100.7 + 201.3 + 201.3 + 100.7 MB, mean 151 - L2 absorbs part of the writes)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 112.7e6


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": p.get("bf16_tflops_sustained", 1402.9), "tflops_burst": p.get("bf16_tflops", 1666.8),
                "hbm_gbs": p.get("hbm_gbs", 6537.6), "source": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = sorted(sm)[len(sm) // 2:]           # upper half = samples under load
        return {"sm_mhz": sorted(hi)[len(hi) // 2], "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- CPU arms
def cpu_pretrain_baseline(batch: int, budget_s: float, threads: int):
    """BASELINE config 1: the reference's forward-surrogate training loop body (pretrain_fwd_model.py:68-92) as the
    oracle port (oracle/models.py: pretrain_step) on the host cores, at the reference's batch size."""
    from oracle import fixtures
    from oracle import models as O
    torch.set_num_threads(threads)
    _, _, f_sd = fixtures.make_weights(42)
    names = [f"model.{i}.{s}" for i in sorted(O.F_LINEAR + O.F_NORM) for s in ("weight", "bias")]
    opt = O.Adam(names, betas=(0.9, 0.999))
    spec, praw, pnorm, mnorm = fixtures.make_batch(batch, seed=12)
    masks = fixtures.make_dropout_masks(batch, seed=5)
    O.pretrain_step(f_sd, opt, pnorm, spec, mnorm, 1e-3, masks)
    t0 = time.perf_counter()
    n = 0
    while True:
        O.pretrain_step(f_sd, opt, pnorm, spec, mnorm, 1e-3, masks)
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 2000:
            break
    return batch * n / el, n, el


def cpu_wide_pretrain_baseline(batch: int, budget_s: float, threads: int, max_steps: int = 50):
    """The same loop body (oracle/models.py: pretrain_step, width-agnostic) at the BASELINE config-5 widths on the
    host cores: (samples/s, steps, seconds)."""
    import numpy as np
    from oracle import models as O
    torch.set_num_threads(threads)
    f_sd = O.init_forward_model(4, WIDE_S, WIDE_MT, WIDE_HIDDEN, gen=torch.Generator().manual_seed(3))
    names = [f"model.{i}.{s}" for i in sorted(O.F_LINEAR + O.F_NORM) for s in ("weight", "bias")]
    opt = O.Adam(names, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(4)
    pnorm = torch.rand(batch, 4, generator=g) * 2 - 1
    spec = -3.0 * torch.rand(batch, WIDE_S, generator=g)
    mnorm = torch.rand(batch, WIDE_MT, generator=g)
    rng = np.random.Generator(np.random.PCG64(5))
    masks = [torch.from_numpy((rng.random((batch, h)) >= 0.2).astype(np.float32)) for h in WIDE_HIDDEN]
    O.pretrain_step(f_sd, opt, pnorm, spec, mnorm, 1e-3, masks)
    t0 = time.perf_counter()
    n = 0
    while True:
        O.pretrain_step(f_sd, opt, pnorm, spec, mnorm, 1e-3, masks)
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= max_steps:
            break
    return batch * n / el, n, el


def run_reference_wide(args):
    """--impl reference --config wide: the oracle port of the surrogate's training loop body at the config-5 widths."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    sample = 1024
    v, n, el = cpu_wide_train_baseline(sample, 1e9, cores, max_steps=max(1, min(args.steps, 10)))
    out = {"impl": "reference", "metric": "widened PI-GAN train samples/s", "value": v, "unit": "samples/s",
           "n_gpus": args.gpus, "steps": n, "warmup": 1, "ms_per_step": el / n * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "PI-GAN train step at the BASELINE config-5 widths", "sample_batch": sample},
           "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                            "sample": f"{n} steps of batch {sample} (oracle/models.py train_step at 2048-wide layers, "
                                      f"fp32, torch CPU)"},
           "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def cpu_train_baseline(batch: int, budget_s: float, threads: int):
    """The oracle port of the reference train step (oracle/models.py, pinned to the reference by tests/golden)
    on the host cores: returns (samples/s, steps timed)."""
    from oracle import fixtures
    from oracle import models as O
    torch.set_num_threads(threads)
    g_sd, d_sd, f_sd = fixtures.make_weights(42)
    og, od = O.Adam(O.G_TRAINABLE), O.Adam(O.D_TRAINABLE)
    spec, praw, pnorm, mnorm = fixtures.make_batch(batch, seed=11)
    b = (spec, praw, pnorm, None, mnorm)
    O.train_step(g_sd, d_sd, f_sd, og, od, b, 2e-4, 2e-4)        # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        O.train_step(g_sd, d_sd, f_sd, og, od, b, 2e-4, 2e-4)
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 200:
            break
    return batch * n / el, n, el


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port), all host threads, bounded sample per step."""
    from oracle import fixtures
    from oracle import models as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = 4096
    g_sd, d_sd, f_sd = fixtures.make_weights(42)
    og, od = O.Adam(O.G_TRAINABLE), O.Adam(O.D_TRAINABLE)
    spec, praw, pnorm, mnorm = fixtures.make_batch(sample, seed=11)
    b = (spec, praw, pnorm, None, mnorm)
    steps = max(1, min(args.steps, 20))
    warm = max(3, args.warmup)            # same warm-up count as the native arm
    for _ in range(warm):
        O.train_step(g_sd, d_sd, f_sd, og, od, b, 2e-4, 2e-4)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(g_sd, d_sd, f_sd, og, od, b, 2e-4, 2e-4)
    el = time.perf_counter() - t0
    v = sample * steps / el
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": el / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PI-GAN train step (D-step + G-step, frozen surrogate), batch {args.batch} per GPU, "
                               "S=250, reference widths", "sample_batch": sample,
                   "extrapolation": f"a CPU step is timed at batch {sample} (bounded sample); samples/s is per-sample "
                                    f"throughput and does not grow with the batch on the CPU (measured 68 vs 78 ms per "
                                    f"4096 rows for the port and the real train_pigan), so it is compared as is with "
                                    f"the GPU's samples/s at batch {args.batch}"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of batch {sample} (oracle/models.py train_step, fp32, torch CPU)"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def bind_to_gpu_numa_node(device_index: int):
    """Pin this rank to the CPU cores next to its GPU (NVML's affinity mask) BEFORE it allocates pinned host
    buffers, so the e2e leg's H2D traffic of the 8 ranks does not cross the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


# ---------------------------------------------------------------------------------------------- widened surrogate
WIDE_S, WIDE_MT, WIDE_HIDDEN = 2048, 8, (2048, 2048, 2048, 2048, 2048)   # BASELINE config 5 widths


def wide_flop_per_sample(S=WIDE_S, Mt=WIDE_MT, hidden=WIDE_HIDDEN) -> int:
    """Forward + dX + dW of the Linear layers (MACs x 2); the first layer has no dX."""
    dims = [4, *hidden, S + Mt]
    macs = [dims[i] * dims[i + 1] for i in range(6)]
    return 2 * (3 * sum(macs) - macs[0])


def e2e_from_host(dev, world, sets, step_fn, KE, barrier):
    """The step fed from pinned host copies of `sets` (tuples of device tensors): H2D of every input on a copy stream,
    double buffered, inside the timed region; the step's losses are read back (D2H) every step.  Returns
    (ms per step as the max over ranks, H2D bytes per step, steps timed)."""
    import torch.distributed as dist
    NS = len(sets)
    host = [tuple(x.cpu().pin_memory() for x in s_) for s_ in sets]
    devb = [tuple(torch.empty_like(x) for x in s_) for s_ in sets]
    copy_stream = torch.cuda.Stream(device=dev)
    h2d_bytes = sum(x.numel() * x.element_size() for x in host[0])

    def h2d(i):
        with torch.cuda.stream(copy_stream):
            for d_, h_ in zip(devb[i % NS], host[i % NS]):
                d_.copy_(h_, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    for rep_ in range(2):   # the second pass is the timed one
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        ev = h2d(0)
        done = [None] * KE
        for i in range(KE):
            torch.cuda.current_stream().wait_event(ev)
            out = step_fn(devb[i % NS])
            done[i] = torch.cuda.Event()
            done[i].record()
            if i + 1 < KE:
                if i + 1 >= NS:
                    copy_stream.wait_event(done[i + 1 - NS])   # the buffer's previous reader has finished
                ev = h2d(i + 1)
            out.tolist()     # D2H of the losses every step
        x1.record()
        barrier()
    te = torch.tensor([x0.elapsed_time(x1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    return float(te.item()) / KE, h2d_bytes, KE


def wide_gan_flop_per_sample(S=WIDE_S, P=4, Mt=WIDE_MT, gh=(2048, 2048), dh=(2048, 2048), fh=WIDE_HIDDEN) -> int:
    """Minimal necessary FLOPs of one PI-GAN step per sample, the formula behind FLOP_PER_TRAIN_SAMPLE:
    G forward once (the two forwards of the reference see the same weights and batch) + backward; D forward on real and
    fake rows + dW + dX of its upper layers (D-step); D forward + dX down to the 4 parameter inputs (G-step); surrogate
    forward (no gradient flows through it: train_pigan.py:156-157)."""
    g1, g2, g3 = S * gh[0], gh[0] * gh[1], gh[1] * P
    d1, d2, d3 = (S + P) * dh[0], dh[0] * dh[1], dh[1]
    G, D = g1 + g2 + g3, d1 + d2 + d3
    dims = [P, *fh, S + Mt]
    F = sum(dims[i] * dims[i + 1] for i in range(6))
    macs = (2 * G + g2 + g3) + (4 * D + 2 * (d2 + d3)) + (D + d3 + d2 + P * dh[0]) + F
    return 2 * macs


def wide_gan_block(dev, world, rank, B, K, W, peaks, barrier, with_e2e: bool):
    """PI-GAN train step at the BASELINE config-5 widths: generator 2048 -> 2048 -> 2048 -> 4, discriminator
    2052 -> 2048 -> 2048 -> 1, frozen surrogate 4 -> 2048 x 5 -> 2056, 2048-point spectra; data parallel = the engine's
    seven phases with NCCL all-reduces of the BatchNorm sums, the two 34 MB gradients and the loss sums between them."""
    import torch.distributed as dist
    from core.models.discriminator import Discriminator
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    from pigan_b200 import engine as E
    from pigan_b200.trainer import NativeTrainer
    torch.manual_seed(11)
    G = Generator(WIDE_S, 4, hidden=(2048, 2048))
    D = Discriminator(WIDE_S, 4, hidden=(2048, 2048))
    F = ForwardModel(4, WIDE_S, WIDE_MT, hidden=WIDE_HIDDEN)
    F.eval()
    tr = NativeTrainer(G, D, F, dev, max_batch=B)
    g = torch.Generator(device=dev)
    g.manual_seed(200 + rank)
    NS = 2   # 2 x 537 MB of spectra > 126 MB L2
    sets = [(-3.0 * torch.rand(B, WIDE_S, device=dev, generator=g),
             2.2 + 0.6 * torch.rand(B, 4, device=dev, generator=g),
             torch.rand(B, WIDE_MT, device=dev, generator=g)) for _ in range(NS)]
    lr = 2e-4
    for i in range(W):
        tr.step(*sets[i % NS], lr, lr)
    barrier()
    l0 = E.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        tr.step(*sets[i % NS], lr, lr)
    e1.record()
    barrier()
    launches = E.launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    losses = tr.losses.tolist()
    tr.engine.profile_begin(None)   # every section, one stream
    for i in range(K):
        tr.step(*sets[i % NS], lr, lr)
    barrier()
    prof = tr.engine.profile_end()
    sections = {k: round(v[1] / K, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    gemm_ms = sum(v for k, v in sections.items() if k.endswith("_gemm"))
    hid_cnt, hid_ms = prof.get("f_hidden_gemm", (0, 0.0))
    flop = wide_gan_flop_per_sample()
    tfl = B * flop / (ms * 1e-3) / 1e12
    info = {"metric": "widened PI-GAN train samples/s", "value": B * world / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": world, "batch_per_gpu": B, "ms_per_step": ms, "steps": K,
            "widths": {"spectrum_dim": WIDE_S, "metrics_dim": WIDE_MT, "generator": [2048, 2048],
                       "discriminator": [2048, 2048], "surrogate": list(WIDE_HIDDEN)},
            "params": {"generator": int(tr.g_grads.numel()), "discriminator": int(tr.d_grads.numel())},
            "exchange": "none (one GPU)" if world == 1 else
            "NCCL all-reduce between the engine's phases: BatchNorm sums (4 x 16 KB), the discriminator's and the "
            "generator's gradients (34 MB each), loss sums",
            "flop_per_sample": flop, "gpu_launches_per_step": launches / K,
            "step_roofline": {"bound": "tensor", "achieved": tfl, "peak": peaks["tflops"], "unit": "TFLOP/s per GPU",
                              "frac": tfl / peaks["tflops"], "frac_of_burst_peak": tfl / peaks["tflops_burst"]},
            "section_ms_per_step_one_stream": sections, "gemm_ms_per_step": gemm_ms,
            "losses_last_step": {"d": losses[0], "g": losses[1], "adv": losses[2]}}
    if hid_cnt:
        ach = 2 * 2048 * 2048 * B / (hid_ms / hid_cnt * 1e-3) / 1e12
        info["roofline"] = {"bound": "tensor", "kernel": "gemm_tc_kernel<EpiStore, row statistics> (frozen surrogate's "
                            "hidden layers 2048 -> 2048; the generator's / discriminator's 2048-wide layers run the same "
                            "kernel family)", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                            "frac": ach / peaks["tflops"], "frac_of_burst_peak": ach / peaks["tflops_burst"],
                            "traffic": None, "launches_timed": hid_cnt, "share_of_step": hid_ms / K / ms}
    if with_e2e:
        ems, h2d_bytes, KE = e2e_from_host(dev, world, sets, lambda t_: tr.step(*t_, lr, lr), max(3, K // 2), barrier)
        info["e2e"] = {"value": B * world / (ems * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                       "d2h_bytes_per_step": 36, "ms_per_step": ems, "steps": KE,
                       "api": "NativeTrainer.step(spectrum, params_denorm, metrics_norm) on batches copied from pinned "
                              "host tensors (copy stream, double buffered), 9 losses read back every step",
                       "h2d_gbs": h2d_bytes / (ems * 1e-3) / 1e9}
    del sets, tr
    torch.cuda.empty_cache()
    return info


def cpu_wide_train_baseline(batch: int, budget_s: float, threads: int, max_steps: int = 20):
    """oracle/models.py train_step (width-agnostic) at the config-5 widths on the host cores."""
    from oracle import models as O
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(3)
    g_sd = O.init_generator(WIDE_S, 4, (2048, 2048), gen)
    d_sd = O.init_discriminator(WIDE_S, 4, (2048, 2048), gen)
    f_sd = O.init_forward_model(4, WIDE_S, WIDE_MT, WIDE_HIDDEN, gen=gen)
    og, od = O.Adam(O.G_TRAINABLE), O.Adam(O.D_TRAINABLE)
    spec = -3.0 * torch.rand(batch, WIDE_S, generator=gen)
    praw = 2.2 + 0.6 * torch.rand(batch, 4, generator=gen)
    mnorm = torch.rand(batch, WIDE_MT, generator=gen)
    b = (spec, praw, (praw - 2.2) / 0.6 * 2 - 1, None, mnorm)
    O.train_step(g_sd, d_sd, f_sd, og, od, b, 2e-4, 2e-4)
    t0 = time.perf_counter()
    n = 0
    while True:
        O.train_step(g_sd, d_sd, f_sd, og, od, b, 2e-4, 2e-4)
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= max_steps:
            break
    return batch * n / el, n, el


def wide_surrogate_block(dev, world, rank, B, K, W, peaks, barrier, with_e2e: bool):
    """Training step of the surrogate at the config-5 widths (4 -> 2048 x 5 -> 2048 + 8) on a surrogate-only engine:
    replicas + NCCL all-reduce of the 84 MB gradient under data parallelism (fwd_trainer.ForwardTrainer)."""
    import torch.distributed as dist
    from core.models.forward_model import ForwardModel
    from pigan_b200 import engine as E
    from pigan_b200.fwd_trainer import ForwardTrainer
    torch.manual_seed(7)
    Fw = ForwardModel(4, WIDE_S, WIDE_MT, hidden=WIDE_HIDDEN)
    ftr = ForwardTrainer(Fw, dev, max_batch=B)
    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)
    NS = 2   # 2 x 537 MB of spectra > 126 MB L2
    sets = [(torch.rand(B, 4, device=dev, generator=g) * 2 - 1,
             -3.0 * torch.rand(B, WIDE_S, device=dev, generator=g),
             torch.rand(B, WIDE_MT, device=dev, generator=g)) for _ in range(NS)]
    for i in range(W):
        ftr.step(*sets[i % NS], 1e-3)
    barrier()
    l0 = E.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        ftr.step(*sets[i % NS], 1e-3)
    e1.record()
    barrier()
    launches = E.launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    # the hidden-layer GEMMs (forward, dX, dW of the four 2048 x 2048 layers) timed on their own stream sections
    ftr.engine.profile_begin(["f_hidden_gemm", "f_dgrad_gemm", "f_wgrad_gemm", "f_out_gemm"])
    for i in range(K):
        ftr.step(*sets[i % NS], 1e-3)
    barrier()
    prof = ftr.engine.profile_end()
    gemm_ms = sum(v[1] for v in prof.values()) / K
    hid_cnt, hid_ms = prof.get("f_hidden_gemm", (0, 0.0))
    flop = wide_flop_per_sample()
    tfl = B * flop / (ms * 1e-3) / 1e12
    info = {"metric": "widened forward-surrogate train samples/s", "value": B * world / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "batch_per_gpu": B, "ms_per_step": ms, "steps": K,
            "widths": {"spectrum_dim": WIDE_S, "metrics_dim": WIDE_MT, "hidden": list(WIDE_HIDDEN)},
            "params": int(ftr.grads.numel()), "gradient_exchange": "none (one GPU)" if world == 1 else
            f"NCCL all-reduce of {ftr.grads.numel() * 4 / 1e6:.0f} MB per step",
            "flop_per_sample": flop, "gpu_launches_per_step": launches / K,
            "step_roofline": {"bound": "tensor", "achieved": tfl, "peak": peaks["tflops"], "unit": "TFLOP/s per GPU",
                              "frac": tfl / peaks["tflops"], "frac_of_burst_peak": tfl / peaks["tflops_burst"]},
            "gemm_ms_per_step": gemm_ms, "gemm_share_of_step": gemm_ms / ms if ms else None,
            "loss_last_step": float(ftr.losses[0])}
    if hid_cnt:
        ach = 2 * 2048 * 2048 * B / (hid_ms / hid_cnt * 1e-3) / 1e12
        info["roofline"] = {"bound": "tensor", "kernel": "gemm_tc_kernel<EpiStore, row statistics> (forward hidden "
                            "layers 2048 -> 2048: Linear + bias + LayerNorm row partials)", "achieved": ach,
                            "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"],
                            "frac_of_burst_peak": ach / peaks["tflops_burst"], "traffic": None,
                            "launches_timed": hid_cnt, "share_of_step": hid_ms / K / ms}
    if with_e2e:
        ems, h2d_bytes, KE = e2e_from_host(dev, world, sets, lambda t: ftr.step(*t, 1e-3), max(3, K // 2), barrier)
        info["e2e"] = {"value": B * world / (ems * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                       "d2h_bytes_per_step": 12, "ms_per_step": ems, "steps": KE,
                       "api": "ForwardTrainer.step on batches copied from pinned host tensors (copy stream, double "
                              "buffered), 3 losses read back every step",
                       "h2d_gbs": h2d_bytes / (ems * 1e-3) / 1e9}
    del sets, ftr
    torch.cuda.empty_cache()
    return info


def run_wide(args):
    """--config wide: the widened surrogate's training step as the line (BASELINE config 5 widths)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        bind_to_gpu_numa_node(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = read_peaks()
    K, W = args.steps, max(3, args.warmup)
    clocks = ClockSampler(local)
    clocks.start()
    info = wide_gan_block(dev, world, rank, args.batch, K, W, peaks, barrier, True)
    clk = clocks.stop()
    sur = wide_surrogate_block(dev, world, rank, args.batch, max(3, min(K, 10)), 3, peaks, barrier, world == 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, n, el = cpu_wide_train_baseline(1024, 15.0, cores)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{n} steps of batch 1024 in {el:.1f} s (oracle/models.py train_step at the same widths, fp32, "
                         f"torch CPU)"}
        v, n, el = cpu_wide_pretrain_baseline(1024, 8.0, cores)
        sur["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                               "sample": f"{n} steps of batch 1024 in {el:.1f} s (oracle/models.py pretrain_step)"}
    if rank == 0:
        out = {"metric": info["metric"], "value": info["value"], "unit": "samples/s", "n_gpus": world,
               "steps": K, "warmup": W, "ms_per_step": info["ms_per_step"],
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f16 operands, f32 accumulate/master", "data": "synthetic",
               "config": {"workload": f"PI-GAN train step (D-step + G-step, frozen surrogate, 7 losses, clip+Adam) at the "
                                      f"BASELINE config-5 widths: generator 2048-2048-2048-4, discriminator "
                                      f"2052-2048-2048-1, surrogate 4-2048x5-2056, S=2048, batch {args.batch} per GPU",
                          "global_batch": args.batch * world, "parallelism": f"dp{world}",
                          "exchange": info["exchange"],
                          "l2": "2 distinct input batches rotated (1.07 GB > 126 MB L2)"},
               "roofline": info.get("roofline"), "step_roofline": info["step_roofline"], "cpu_baseline": cpu,
               "e2e": info.get("e2e"), "gpu_launches": int(info["gpu_launches_per_step"] * K),
               "clocks": clk, "wide_pigan_training": info, "wide_surrogate_training": sur}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- native arm
def run_native(args):
    import torch.distributed as dist
    from core.models.discriminator import Discriminator
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    from pigan_b200 import engine as E
    from pigan_b200 import flat, scoring, synthetic
    from pigan_b200.trainer import NativeTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = 0
    if world > 1:
        numa_cpus = bind_to_gpu_numa_node(local)   # N=1 keeps all cores (the CPU baseline runs on them)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, max(3, args.warmup)
    peaks = read_peaks()

    torch.manual_seed(42)
    G, D, F = Generator(250, 4), Discriminator(250, 4), ForwardModel(4, 250, 8)
    F.eval()
    tr = NativeTrainer(G, D, F, dev, max_batch=B)
    # inputs larger than L2: NSETS distinct batches (NSETS * 68.7 MB at B=65536 > 126 MB L2), rotated per step
    NSETS = 4
    sets = []
    for i in range(NSETS):
        sp, pr, _pn, mn = synthetic.make_batch(B, 250, seed=1000 * rank + i, device=dev)
        sets.append((sp, pr, mn))
    lr = 2e-4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        tr.step(*sets[i % NSETS], lr, lr)
    barrier()

    # ---- timed region 1: device-resident inputs, EXACTLY K steps in the product configuration (the frozen surrogate's
    # forward chain of the G-step runs on the engine's second stream beside the D-step)
    DOMINANT = "f_hidden_gemm"
    clocks = ClockSampler(local)
    clocks.start()
    l0 = E.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        tr.step(*sets[i % NSETS], lr, lr)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = E.launch_count() - l0
    # ---- timed region 1b: the same K steps again with the dominant kernel bracketed by CUDA events on its launch
    # stream (pigan_engine_profile_begin/_end).  The engine keeps ONE stream while it profiles, so the kernel's launches
    # are not time-shared with the D-step's kernels: this region gives the kernel's own duration for the roofline
    # block, region 1 the step time.
    tr.engine.profile_begin([DOMINANT])
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for i in range(K):
        tr.step(*sets[i % NSETS], lr, lr)
    p1.record()
    barrier()
    ms_prof = p0.elapsed_time(p1)
    prof = tr.engine.profile_end()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * K / (ms * 1e-3)
    losses = tr.losses.tolist()

    # dominant kernel: the 4 hidden-layer GEMMs of the frozen surrogate (256->512->1024->512->256), 1 245 184 MAC/row
    dom_cnt, dom_ms = prof.get(DOMINANT, (0, 0.0))
    dom_flop_per_launch = 2 * (256 * 512 + 512 * 1024 + 1024 * 512 + 512 * 256) * B / 4.0
    roof = None
    if dom_cnt:
        ach = dom_flop_per_launch / (dom_ms / dom_cnt * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel<EpiLnStore> (forward-surrogate hidden layers: Linear+LayerNorm+LeakyReLU)",
                "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"],
                "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "traffic_source": "profiles/r02_gemm_kernels_ncu_full.csv "
                "(ncu --set full, dram__bytes_read+write, mean of the 4 EpiLnStore launches)", "peak_source": peaks["source"] + " bf16 sustained",
                "frac_of_burst_peak": ach / peaks["tflops_burst"], "peak_burst": peaks["tflops_burst"],
                "timed_region_ms": ms_prof, "note": "kernel durations from CUDA events on the launch stream over a second "
                "timed region of the same K steps in which the engine keeps one stream (in the product configuration "
                "the surrogate chain shares the SMs with the D-step on a second stream); the region is tens of "
                "milliseconds at full clocks: between the burst and the sustained (seconds-long, power-capped) cuBLAS "
                "figure; both fractions are given",
                "launches_timed": dom_cnt, "share_of_step": dom_ms / ms_prof,
                "ms_per_step_one_stream": ms_prof / K}
    step_tflops = FLOP_PER_TRAIN_SAMPLE * B * K / (ms * 1e-3) / 1e12

    # ---- data parallel: what each exchange point of the step costs on this rank (kernel + wait for the slowest peer),
    # CUDA events around the exchange kernels in a few extra, untimed steps
    dp_exchange_us = None
    if world > 1 and tr.xchg is not None:
        tr.exchange_events = []
        for i in range(5):
            tr.step(*sets[i % NSETS], lr, lr)
        torch.cuda.synchronize()
        acc = {}
        for name, a, b in tr.exchange_events:
            acc.setdefault(name, []).append(a.elapsed_time(b) * 1e3)
        tr.exchange_events = None
        dp_exchange_us = {k: round(sum(v) / len(v), 1) for k, v in acc.items()}
        barrier()

    # ---- the same step at a batch that tiles the 148 SMs without a partial wave (592 row tiles = 4 per SM instead of
    # 512 = 3.46): informational, shows how much of the gap to the roofline is wave quantisation at B = 65536
    quant = None
    if world == 1 and not args.no_quant_probe:
        Bq = 148 * 128 * 4
        del tr
        torch.cuda.empty_cache()
        torch.manual_seed(42)
        Gq, Dq, Fq = Generator(250, 4), Discriminator(250, 4), ForwardModel(4, 250, 8)
        Fq.eval()
        trq = NativeTrainer(Gq, Dq, Fq, dev, max_batch=Bq)
        qs = [synthetic.make_batch(Bq, 250, seed=77 + i, device=dev) for i in range(3)]
        for i in range(3):
            trq.step(qs[i % 3][0], qs[i % 3][1], qs[i % 3][3], lr, lr)
        barrier()
        e0.record()
        for i in range(K):
            trq.step(qs[i % 3][0], qs[i % 3][1], qs[i % 3][3], lr, lr)
        e1.record()
        barrier()
        msq = e0.elapsed_time(e1) / K
        quant = {"batch": Bq, "ms_per_step": msq, "value": Bq / (msq * 1e-3), "unit": "samples/s",
                 "frac_of_step_roofline": FLOP_PER_TRAIN_SAMPLE * Bq / (msq * 1e-3) / 1e12 / peaks["tflops"]}
        del trq, qs, Gq, Dq, Fq
        torch.cuda.empty_cache()
        torch.manual_seed(42)
        G, D, F = Generator(250, 4), Discriminator(250, 4), ForwardModel(4, 250, 8)
        F.eval()
        tr = NativeTrainer(G, D, F, dev, max_batch=B)
        for i in range(3):
            tr.step(*sets[i % NSETS], lr, lr)
        barrier()

    # ---- the same step with the physics-metric loss term switched on (lambda = 0 is the reference):
    # F forward with the fp32 dump, physics metrics forward + backward, surrogate VJP, then the fused phases
    pm_info = None
    if world == 1 and not args.no_quant_probe:
        torch.manual_seed(42)
        Gp, Dp, Fp = Generator(250, 4), Discriminator(250, 4), ForwardModel(4, 250, 8)
        Fp.eval()
        trp = NativeTrainer(Gp, Dp, Fp, dev, max_batch=B, lambda_physics_metric=1.0)
        for i in range(3):
            trp.step(*sets[i % NSETS], lr, lr)
        barrier()
        e0.record()
        for i in range(max(3, K // 2)):
            trp.step(*sets[i % NSETS], lr, lr)
        e1.record()
        barrier()
        mspm = e0.elapsed_time(e1) / max(3, K // 2)
        pm_info = {"ms_per_step": mspm, "extra_ms": mspm - ms / K, "value": B / (mspm * 1e-3), "unit": "samples/s",
                   "physics_metric_loss": float(trp.last_physics_metric_loss),
                   "rows_with_defined_metrics": int(trp.last_physics_metric_rows),
                   "note": "random-initialised surrogate: its reconstructions rarely have a half-depth crossing, so few "
                           "or no rows contribute to the loss VALUE here; the cost of the term is the same either way",
                   "what": "NativeTrainer(lambda_physics_metric=1): + F forward (fp32 dump), peak metrics fwd/bwd of "
                           "reconstruction and target, surrogate VJP (pigan_forward_model_vjp), one stream"}
        del trp, Gp, Dp, Fp
        torch.cuda.empty_cache()
        tr.engine.load_forward_model(tr.fs.params.tensor())

    # ---- timed region 2: end to end from pinned HOST buffers through the public API.  Two variants:
    #   e2e      NativeTrainer.step_prepared: the dataset keeps the fp16 first-layer operand that
    #            NativeTrainer.prepare_operand built once (the analogue of the reference's one-time dataset
    #            normalisation) -> 35.7 MB per step cross PCIe
    #   e2e_fp32 NativeTrainer.step on the raw fp32 tensors the reference's DataLoader yields -> 68.7 MB per step
    center = torch.cat([s_[0][:512] for s_ in sets]).mean(dim=0).contiguous()
    if world > 1:
        dist.all_reduce(center)
        center /= world
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()

    def run_e2e(host_sets, step_fn, n_steps):
        """H2D on a copy stream into a ring of NBUF device buffers (prefetch depth NBUF - 1), one step per batch,
        D2H of the 9 losses every step"""
        NBUF = 3
        dbuf = [tuple(torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host_sets[0]) for _ in range(NBUF)]
        loss_host = torch.empty(n_steps, 9, dtype=torch.float32).pin_memory()
        ready = [torch.cuda.Event() for _ in range(NBUF)]
        freed = [torch.cuda.Event() for _ in range(NBUF)]

        def h2d(i):
            slot = i % NBUF
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[slot])
                for d, h in zip(dbuf[slot], host_sets[i % NSETS]):
                    d.copy_(h, non_blocking=True)
                ready[slot].record(copy_stream)

        def loop(n):
            for s_ in range(NBUF):
                freed[s_].record(main)
            for j in range(min(NBUF - 1, n)):
                h2d(j)
            for i in range(n):
                if i + NBUF - 1 < n:
                    h2d(i + NBUF - 1)
                slot = i % NBUF
                main.wait_event(ready[slot])
                ls = step_fn(dbuf[slot])
                freed[slot].record(main)
                loss_host[i].copy_(ls, non_blocking=True)

        loop(min(3, n_steps))
        barrier()
        e0.record()
        loop(n_steps)
        e1.record()
        barrier()
        t_ = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item()), sum(h.numel() * h.element_size() for h in host_sets[0])

    host_prep = []
    for sp, pr, mn in sets:
        op = NativeTrainer.prepare_operand(sp, pr, center)
        host_prep.append((op.cpu().pin_memory(), mn.cpu().pin_memory()))
    ms2, h2d_bytes_prep = run_e2e(host_prep, lambda b: tr.step_prepared(b[0], center, b[1], lr, lr), K)
    e2e_prepared = {"value": B * world * K / (ms2 * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes_prep,
                    "d2h_bytes_per_step": 36, "ms_per_step": ms2 / K,
                    "api": "NativeTrainer.step_prepared(fp16 operand prepared once per dataset, metrics_norm)"}
    del host_prep
    host_raw = [tuple(x.cpu().pin_memory() for x in s_) for s_ in sets]
    ms2b, h2d_bytes = run_e2e(host_raw, lambda b: tr.step(b[0], b[1], b[2], lr, lr), K)
    e2e = {"value": B * world * K / (ms2b * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
           "d2h_bytes_per_step": 36, "ms_per_step": ms2b / K,
           "api": "NativeTrainer.step(spectrum, params_denorm, metrics_norm) on the fp32 host tensors the reference's "
                  "DataLoader yields (pinned), H2D on a copy stream, 9 losses read back every step",
           "critical_path": "h2d" if ms2b / K > 1.15 * ms / K else "step",
           "h2d_gbs": h2d_bytes * world / (ms2b / K * 1e-3) / 1e9}
    del host_raw

    # ---- inverse-design scoring (BASELINE config 4): candidates/s, sharded by candidate, final top-k gather
    G.eval()
    gs = flat.net_state(G, "generator")
    chunk = scoring.full_wave_chunk(dev)      # 148 SMs x 128 rows x 4 waves = 75 776: no partial wave per GEMM
    seng = E.Engine(chunk, dev)
    seng.load_forward_model(tr.fs.params.tensor())
    designer = scoring.InverseDesigner(seng, gs.params.tensor(), gs.bn.tensor(), chunk=chunk)
    target = sets[0][0][0].clone()
    n_cand = args.candidates               # TOTAL over all ranks (BASELINE config 4: 100 M, strong scaling)
    designer.search(target, 4 * chunk * world, k=1024)
    barrier()
    e0.record()
    res = designer.search(target, n_cand, k=1024)
    e1.record()
    barrier()
    ms3 = e0.elapsed_time(e1)
    t = torch.tensor([ms3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms3 = float(t.item())
    cand_s = n_cand / (ms3 * 1e-3)
    clk = clocks.stop()     # sampled across the three timed regions (device-resident, end-to-end, scoring)
    score_tf = cand_s * FLOP_PER_CANDIDATE / 1e12 / world
    score_info = {"metric": "inverse-design candidates/s", "value": cand_s, "unit": "candidates/s",
                  "candidates": n_cand, "candidates_per_gpu": n_cand // world, "scaling": "strong", "n_gpus": world,
                  "per_gpu": cand_s / world, "k": 1024, "ms": ms3,
                  "best_recon_error": float(res["recon_error"][0]),
                  "roofline": {"bound": "tensor", "achieved": score_tf, "peak": peaks["tflops"],
                               "unit": "TFLOP/s per GPU", "frac": score_tf / peaks["tflops"],
                               "frac_of_burst_peak": score_tf / peaks["tflops_burst"],
                               "flop_per_candidate": FLOP_PER_CANDIDATE,
                               "target_50pct": 0.5 * peaks["tflops"] * 1e12 / FLOP_PER_CANDIDATE},
                  "tensor_frac": score_tf / peaks["tflops"],
                  "chunk": chunk, "collective": "one all_gather of k rows per rank + local merge" if world > 1 else "none",
                  "noise": "in-kernel Philox4x32-10 keyed by (seed, global candidate index)"}
    # model-validation loop (unified_evaluator.py:439-468): cycle error + stability + plausibility per row, one call
    vn = min(chunk, 65536)
    vspec = sets[0][0][:vn]
    vnoise = torch.randn(vn, 250, device=dev)
    for _ in range(2):
        seng.validate(gs.params.tensor(), gs.bn.tensor(), vspec, vnoise, 0.01)
    barrier()
    e0.record()
    for _ in range(5):
        vres = seng.validate(gs.params.tensor(), gs.bn.tensor(), vspec, vnoise, 0.01)
    e1.record()
    barrier()
    msv = e0.elapsed_time(e1) / 5
    score_info["model_validation"] = {"rows_per_s": vn / (msv * 1e-3), "rows": vn, "ms": msv,
                                      "cycle_error_mean": float(vres["cycle_error"].mean()),
                                      "stability_mean": float(vres["stability"].mean()),
                                      "plausibility_mean": float(vres["plausibility"].mean())}
    del vspec, vnoise
    del designer, seng

    # ---- physics metrics (BASELINE config 3): sweep 1 M .. 64 M spectra, HBM-bound, sharded by spectrum
    from pigan_b200 import native
    phys_info = None
    sizes = [int(x) for x in args.physics_sweep.split(",") if x] if args.physics_sweep else []
    if sizes:
        from pigan_b200 import physics as phys_mod
        freq = synthetic.frequencies(250, device=dev)
        sweep = []
        lo, hi = phys_mod.shard_rows(max(sizes), rank, world)       # sharded by spectrum, no collective
        base = sets[0][0]
        try:
            spec_big = base.repeat((hi - lo + B - 1) // B, 1)[: hi - lo].contiguous()
        except torch.OutOfMemoryError:
            spec_big = None
            sizes = [n for n in sizes if n * 1000 < 40e9]
            lo, hi = phys_mod.shard_rows(max(sizes), rank, world)
            spec_big = base.repeat((hi - lo + B - 1) // B, 1)[: hi - lo].contiguous()
        o_idx = torch.empty(hi - lo, device=dev, dtype=torch.int32)
        o_met = torch.empty(hi - lo, 4, device=dev, dtype=torch.float32)
        for n_total in sizes:
            a, b_ = phys_mod.shard_rows(n_total, rank, world)
            n_loc = b_ - a

            def phys():
                native.check(native.lib.pigan_physics_metrics(spec_big.data_ptr(), n_loc, 250, freq.data_ptr(), None,
                                                              0.0, o_idx.data_ptr(), o_met.data_ptr(),
                                                              native.current_stream()))
            for _ in range(3):
                phys()
            barrier()
            e0.record()
            for _ in range(5):
                phys()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms4 = float(t.item())
            gbs = n_total * 1016 / (ms4 * 1e-3) / 1e9 / world
            sweep.append({"spectra": n_total, "ms": ms4, "value": n_total / (ms4 * 1e-3), "unit": "spectra/s",
                          "gbs_per_gpu": gbs, "frac": gbs / peaks["hbm_gbs"]})
        best = max(sweep, key=lambda r: r["value"])
        phys_info = {"metric": "physics-metric spectra/s", "value": best["value"], "unit": "spectra/s",
                     "spectra": best["spectra"], "ms": best["ms"], "n_gpus": world, "scaling": "weak-by-size sweep, "
                     "sharded by spectrum over the ranks, no collective",
                     "input": f"the {B}-row synthetic batch tiled to the sweep size (distinct rows do not matter for "
                              f"an HBM-bound row kernel; {best['spectra'] * 1000 / 1e9:.0f} GB >> 126 MB L2)",
                     "sweep": sweep,
                     "roofline": {"bound": "hbm", "achieved": best["gbs_per_gpu"], "peak": peaks["hbm_gbs"],
                                  "unit": "GB/s per GPU", "frac": best["gbs_per_gpu"] / peaks["hbm_gbs"],
                                  "bytes_per_spectrum": 1016}}
        # backward of the same kernel: vector-Jacobian product into the spectra, at the 4 M point
        n_b = min(hi - lo, (1 << 22) // world)
        g_met = torch.ones(n_b, 4, device=dev)
        g_spec = torch.empty(n_b, 250, device=dev)

        def phys_bwd():
            native.check(native.lib.pigan_physics_metrics_backward(spec_big.data_ptr(), n_b, 250, freq.data_ptr(), None,
                                                                   0.0, g_met.data_ptr(), g_spec.data_ptr(), None, None,
                                                                   native.current_stream()))
        for _ in range(2):
            phys_bwd()
        barrier()
        e0.record()
        for _ in range(5):
            phys_bwd()
        e1.record()
        barrier()
        ms4b = e0.elapsed_time(e1) / 5
        gbs_b = n_b * 2016 / (ms4b * 1e-3) / 1e9
        phys_info["backward"] = {"value": n_b / (ms4b * 1e-3), "unit": "spectra/s per GPU", "spectra": n_b, "ms": ms4b,
                                 "roofline": {"bound": "hbm", "achieved": gbs_b, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                              "frac": gbs_b / peaks["hbm_gbs"], "bytes_per_spectrum": 2016}}
        del spec_big, o_idx, o_met, g_met, g_spec
        torch.cuda.empty_cache()

    # ---- on-device data pipeline: synthetic-spectrum generator, shuffled batch gather, and the
    # train step fed from a resident dataset (gather of the fp16 operand + metrics rows, then step_prepared)
    from pigan_b200 import device_data
    n_gen = 1 << 20
    gspec, gpar = torch.empty(n_gen, 250, device=dev), torch.empty(n_gen, 4, device=dev)
    gfreq = synthetic.frequencies(250, device=dev)
    for _ in range(3):
        device_data.generate_spectra(n_gen, dev, seed=1, frequency=gfreq, out=gspec, params_out=gpar)
    barrier()
    e0.record()
    for i in range(10):
        device_data.generate_spectra(n_gen, dev, seed=2 + i, frequency=gfreq, out=gspec, params_out=gpar)
    e1.record()
    barrier()
    ms7 = e0.elapsed_time(e1) / 10
    del gspec, gpar
    n_res = 4 * B
    res_op = torch.cat([NativeTrainer.prepare_operand(s_[0], s_[1], center) for s_ in sets])
    res_mn = torch.cat([s_[2] for s_ in sets])
    gperm = torch.Generator(device=dev)
    gperm.manual_seed(5)
    perm = torch.randperm(n_res, generator=gperm, device=dev)
    ob, mb = torch.empty(B, 256, device=dev, dtype=torch.float16), torch.empty(B, 8, device=dev)
    for i in range(3):
        idx = perm[(i % 4) * B:(i % 4 + 1) * B]
        device_data.gather_rows(res_op, idx, ob); device_data.gather_rows(res_mn, idx, mb)
        tr.step_prepared(ob, center, mb, lr, lr)
    barrier()
    e0.record()
    for i in range(K):
        idx = perm[(i % 4) * B:(i % 4 + 1) * B]
        device_data.gather_rows(res_op, idx, ob); device_data.gather_rows(res_mn, idx, mb)
        tr.step_prepared(ob, center, mb, lr, lr)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms8 = float(t.item()) / K
    e0.record()
    for i in range(20):
        device_data.gather_rows(res_op, perm[(i % 4) * B:(i % 4 + 1) * B], ob)
    e1.record()
    barrier()
    ms9 = e0.elapsed_time(e1) / 20
    gen_gbs = n_gen * 1016 / (ms7 * 1e-3) / 1e9
    gat_gbs = 2 * B * 512 / (ms9 * 1e-3) / 1e9
    pipe_info = {"generator": {"value": n_gen / (ms7 * 1e-3), "unit": "spectra/s", "rows": n_gen, "ms": ms7,
                               "roofline": {"bound": "hbm", "achieved": gen_gbs, "peak": peaks["hbm_gbs"],
                                            "unit": "GB/s", "frac": gen_gbs / peaks["hbm_gbs"],
                                            "bytes_per_spectrum": 1016,
                                            "note": "issue-bound in practice: 2 expf + Philox/Box-Muller per element"}},
                 "gather": {"rows": B, "row_bytes": 512, "ms": ms9,
                            "roofline": {"bound": "hbm", "achieved": gat_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                         "frac": gat_gbs / peaks["hbm_gbs"]}},
                 "loader_fed_train": {"value": B * world / (ms8 * 1e-3), "unit": "samples/s", "ms_per_step": ms8,
                                      "api": "gather_rows(resident fp16 operand + metrics, shuffled index) + "
                                             "NativeTrainer.step_prepared", "resident_rows": n_res}}
    del res_op, res_mn, ob, mb

    # ---- evaluator reductions: calculate_metrics over [n, 250] result arrays, HBM-bound
    from pigan_b200 import evalstats
    n_ev = 1 << 20
    yt = sets[0][0].repeat(n_ev // B + 1, 1)[:n_ev].contiguous()
    yp = yt + 0.25
    rm = evalstats.RegressionMetrics(250, dev)
    for _ in range(2):
        rm.update(yt, yp)
    barrier()
    e0.record()
    for _ in range(5):
        rm.update(yt, yp)
    e1.record()
    barrier()
    ms6 = e0.elapsed_time(e1) / 5
    ev_gbs = 2 * n_ev * 250 * 4 / (ms6 * 1e-3) / 1e9
    eval_info = {"metric": "evaluator regression-metric rows/s", "value": n_ev / (ms6 * 1e-3), "unit": "rows/s",
                 "rows": n_ev, "cols": 250, "ms": ms6, "mse": rm.compute()["mse"],
                 "roofline": {"bound": "hbm", "achieved": ev_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": ev_gbs / peaks["hbm_gbs"], "bytes_per_row": 2000}}
    del yt, yp, rm

    # ---- forward-surrogate training step (BASELINE config 1 moved to the GPU): replicas + gradient
    # all-reduce under data parallelism.  Runs last: it replaces the engine's frozen-surrogate state.
    from pigan_b200.fwd_trainer import ForwardTrainer
    Ft = ForwardModel(4, 250, 8)
    Ft.load_state_dict({k: v.detach().clone() for k, v in F.state_dict().items()})
    ftr = ForwardTrainer(Ft, dev, max_batch=B, engine=tr.engine)
    pnorm_sets = [((s_[1] - 2.5) / 0.3).contiguous() for s_ in sets]   # data_loader.py:185-196 normalisation
    for i in range(3):
        ftr.step(pnorm_sets[i % NSETS], sets[i % NSETS][0], sets[i % NSETS][2], 1e-3)
    barrier()
    e0.record()
    for i in range(K):
        ftr.step(pnorm_sets[i % NSETS], sets[i % NSETS][0], sets[i % NSETS][2], 1e-3)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms5 = float(t.item()) / K
    f_tflops = B * FLOP_PER_PRETRAIN_SAMPLE / (ms5 * 1e-3) / 1e12
    fwd_info = {"metric": "forward-surrogate train samples/s", "value": B * world / (ms5 * 1e-3), "unit": "samples/s",
                "batch_per_gpu": B, "ms_per_step": ms5, "dropout": "counter-based Philox, p=0.2",
                "tensor_frac": f_tflops / peaks["tflops"], "flop_per_sample": FLOP_PER_PRETRAIN_SAMPLE,
                "loss_last_step": float(ftr.losses[0])}
    del pnorm_sets, ftr

    # ---- widened surrogate (BASELINE config 5 widths: hidden 2048, 2048-point spectra) on its own surrogate-only
    # engine, a few steps (the `--config wide` line times it on its own, with an end-to-end leg)
    wide_info = wide_gan_info = None
    if not args.no_wide:
        # (side blocks of the line: a host-side failure here must not cost the headline numbers above)
        try:
            wide_info = wide_surrogate_block(dev, world, rank, min(B, 65536), max(3, min(K, 8)), 3, peaks, barrier, False)
        except Exception as exc:   # noqa: BLE001
            wide_info = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            wide_gan_info = wide_gan_block(dev, world, rank, min(B, 65536), max(3, min(K, 6)), 3, peaks, barrier, False)
        except Exception as exc:   # noqa: BLE001
            wide_gan_info = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- CPU baseline (rank 0, N=1 only): oracle port of the reference step on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, n, el = cpu_train_baseline(4096, 12.0, cores)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{n} steps of batch 4096 in {el:.1f} s (oracle/models.py train_step, fp32, torch CPU)"}
        v, n, el = cpu_pretrain_baseline(64, 6.0, cores)
        fwd_info["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} steps of batch 64 (cfg.BATCH_SIZE, BASELINE config 1) in {el:.1f} s "
                                              f"(oracle/models.py pretrain_step, fp32, torch CPU)"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate/master", "data": "synthetic",
            "config": {"workload": f"PI-GAN train step (D-step + G-step, frozen surrogate, 7 losses, clip+Adam), "
                                   f"batch {B} per GPU, S=250, reference widths",
                       "global_batch": B * world, "parallelism": f"dp{world}",
                       "exchange": ("none (one GPU)" if world == 1 else
                                    ("peer: one-shot all-reduce kernels over cudaIpc-mapped NVLink memory (csrc/dp.cu)"
                                     if tr.xchg is not None else "nccl all-reduce")),
                       "exchange_us_rank0": dp_exchange_us,
                       "rank_cpu_affinity": f"{numa_cpus} GPU-local cores per rank (NVML)" if numa_cpus else "default",
                       "l2": f"{NSETS} distinct input batches rotated ({NSETS * h2d_bytes / 1e6:.0f} MB > 126 MB L2)"},
            "roofline": roof,
            "step_roofline": {"bound": "tensor", "achieved": step_tflops, "peak": peaks["tflops"],
                              "unit": "TFLOP/s per GPU", "frac": step_tflops / peaks["tflops"],
                              "flop_per_sample": FLOP_PER_TRAIN_SAMPLE},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "e2e_prepared_operand": e2e_prepared,
            "gpu_launches": int(launches),
            "clocks": clk,
            "scoring": score_info,
            "physics": phys_info,
            "surrogate_training": fwd_info,
            "wide_pigan_training": wide_gan_info,
            "wide_surrogate_training": wide_info,
            "evaluator_reductions": eval_info,
            "data_pipeline": pipe_info,
            "wave_quantisation_probe": quant,
            "physics_metric_term": pm_info,
            "losses_last_step": {"d": losses[0], "g": losses[1], "adv": losses[2]},
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--candidates", type=int, default=100_000_000,
                    help="TOTAL candidates of the scoring line (BASELINE config 4: 100 M), split over the ranks")
    ap.add_argument("--physics-sweep", default="1048576,4194304,16777216,67108864",
                    help="total spectra of the physics-kernel sweep (BASELINE config 3), comma separated; '' skips")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="default", choices=["default", "wide"],
                    help="wide: the line is the widened surrogate's training step (BASELINE config 5 widths)")
    ap.add_argument("--no-wide", action="store_true", help="skip the widened-surrogate block of the default line")
    ap.add_argument("--no-quant-probe", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_wide(args) if args.config == "wide" else run_reference(args)
    elif args.config == "wide":
        run_wide(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
