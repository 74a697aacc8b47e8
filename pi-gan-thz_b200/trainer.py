"""Native PI-GAN trainer state: flat parameter / gradient / Adam buffers, the C-ABI argument block and the
data-parallel phase schedule.  Used by the drop-in ``core.train.train_pigan.train_pigan`` and by bench.py.

Reference semantics reproduced (core/train/train_pigan.py:114-187): one D-step then one G-step per batch,
label smoothing 0.9/0.1 (D) and 1.0 (G), frozen forward surrogate, the weighted loss of :174-181,
clip_grad_norm_(1.0), Adam(betas=(0.5, 0.999)); BatchNorm running statistics advance twice per step (F8).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import dp as _dp
from . import engine as _engine
from . import flat as _flat
from . import native

LOSS_KEYS = _engine.LOSS_KEYS


def dp_phase_plan(h1: int, h2: int):
    """[(phase, [(buffer, slice or None)])]: what every rank all-reduces (sum) after each phase of
    pigan_train_step_phase.  h1, h2 = the generator's hidden widths (BatchNorm-1 / BatchNorm-2 channels)."""
    return [
        (0, [("bn_sums", slice(0, 2 * h1))]),                       # BatchNorm-1 sum, sum of squares
        (1, [("bn_sums", slice(2 * h1, 2 * h1 + 2 * h2))]),         # BatchNorm-2
        (2, [("d_grads", None)]),                                   # discriminator gradients (before clip+Adam)
        (3, [("bn_bwd_sums", slice(0, 2 * h2))]),                   # BatchNorm-2 backward: sum dy, sum dy*xhat
        # BatchNorm-1 backward + the loss numerators (complete after phase 3: they ride along, one exchange fewer)
        (4, [("bn_bwd_sums", slice(2 * h2, 2 * h2 + 2 * h1)), ("loss_sums", slice(0, 8))]),
        (5, [("g_grads", None)]),                                   # generator gradients
        (6, []),                                                    # clip+Adam(G), loss finalisation
    ]


def check_equal_batch(batch: int, last_checked, group=None) -> int:
    """Data parallel: all ranks must step with the same number of rows (global batch = rows x world enters BatchNorm
    statistics, loss means and the gradient scale; a rank that stops early leaves its peers spinning in the peer
    exchange kernels).  Called before anything is launched whenever the local batch size changes (the ragged last
    batch of an epoch changes it on all ranks at once); raises on EVERY rank when the sizes differ.  A loader that
    changes the size on some ranks only is not caught here - device_data.DeviceLoader(rank, world) never does.
    Returns the size to remember."""
    if batch == last_checked:
        return last_checked
    world = dist.get_world_size(group)
    sizes = [None] * world
    dist.all_gather_object(sizes, int(batch), group=group)
    if len(set(sizes)) != 1:
        raise RuntimeError(f"data-parallel step with unequal local batches {sizes}: use drop_last / "
                           "device_data.DeviceLoader(rank, world), which cuts ragged batches on all ranks alike")
    if batch < 2:
        raise RuntimeError("data-parallel step needs at least 2 rows per rank")
    return batch


def run_dp_step(run_phase, get_buffer, all_reduce, plan) -> None:
    """One data-parallel step: engine phases in order, each followed by its all-reduces (in place, sum)."""
    for phase, reductions in plan:
        run_phase(phase)
        for name, sl in reductions:
            buf = get_buffer(name)
            all_reduce(buf if sl is None else buf[sl])


class NativeTrainer:
    def __init__(self, generator, discriminator, forward_model, device, max_batch: int, cfg=None,
                 f1_idx: int = 0, f2_idx: int = 1, process_group=None, lambda_physics_metric: float = 0.0,
                 physics_metric_weights=(1.0, 1e-2, 1.0, 1e-2)):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("NativeTrainer needs a CUDA device — the B200 path has no CPU fallback")
        self.g, self.d, self.f = generator, discriminator, forward_model
        for m in (generator, discriminator, forward_model):
            m.to(self.device)
        self.gs = _flat.net_state(generator, "generator")
        self.ds = _flat.net_state(discriminator, "discriminator")
        self.fs = _flat.net_state(forward_model, "forward_model")
        self.engine = _engine.Engine(max_batch, self.device, self._dims_of(generator, discriminator, forward_model))
        self.wide = self.engine.dims is not None and native.dims_key(self.engine.dims) != native.dims_key(native.default_dims())
        self.engine.load_forward_model(self.fs.params.tensor())
        gp, dp = self.gs.params.tensor(), self.ds.params.tensor()
        self.g_grads, self.g_m, self.g_v = (torch.zeros_like(gp) for _ in range(3))
        self.d_grads, self.d_m, self.d_v = (torch.zeros_like(dp) for _ in range(3))
        self.losses = torch.zeros(9, device=self.device, dtype=torch.float32)
        self.step_count = 0
        self._center = None
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        lam = dict(lambda_recon=100.0, lambda_physics_spectrum=10.0, lambda_physics_metrics=1.0, lambda_maxwell=1.0,
                   lambda_lc=1.0, lambda_param_range=0.1, lambda_bnn_kl=0.0)
        if cfg is not None:
            lam = dict(lambda_recon=cfg.LAMBDA_RECON, lambda_physics_spectrum=cfg.LAMBDA_PHYSICS_SPECTRUM,
                       lambda_physics_metrics=cfg.LAMBDA_PHYSICS_METRICS, lambda_maxwell=cfg.LAMBDA_MAXWELL,
                       lambda_lc=cfg.LAMBDA_LC, lambda_param_range=cfg.LAMBDA_PARAM_RANGE,
                       lambda_bnn_kl=cfg.LAMBDA_BNN_KL)
        self.lam = lam
        self.f1_idx, self.f2_idx = int(f1_idx), int(f2_idx)
        # SURVEY 8(f) N2 (the reference has no such term, default 0 = the reference's loss): a physics-metric loss
        #   L_pm = mean_rows sum_k w_k (m_k(F(G(x)).spectrum) - m_k(x))^2,  m = (f_res, Q, FoM, S) at the row's minimum,
        # over rows where both sides have a defined Q; it reaches the generator through the metrics' own backward
        # (pigan_physics_metrics_backward) and the frozen surrogate's vector-Jacobian product (pigan_forward_model_vjp)
        self.lambda_physics_metric = float(lambda_physics_metric)
        self.physics_metric_weights = tuple(float(x) for x in physics_metric_weights)
        self.last_physics_metric_loss = None
        self.exchange_events = None      # set to [] to collect (name, start, end) CUDA events of every peer exchange
        # data parallel: gradients are produced straight into an NVLink-mapped exchange region and summed by the
        # library's one-shot kernels (dp.py); NCCL all-reduce is the fallback when peers cannot be mapped
        # (widened dims: 8.4 M-float gradients are bandwidth-, not latency-bound - NCCL's all-reduce, not the one-shot
        # peer kernels that read every peer's whole buffer)
        self.xchg = _dp.DpExchange.create(max(gp.numel(), dp.numel()), self.device, process_group) \
            if (self.world > 1 and not self.wide) else None
        if self.xchg is not None:
            self._slots = {(net, par): self.xchg.grad_slot(net, par, n)
                           for net, n in ((0, gp.numel()), (1, dp.numel())) for par in (0, 1)}
        # expose gradients on the parameters the way autograd would (views of the flat buffers)
        for p, gview in zip(self.gs.params._tensors(), self.gs.params.views_like(self.g_grads)):
            p.grad = gview
        for p, gview in zip(self.ds.params._tensors(), self.ds.params.views_like(self.d_grads)):
            p.grad = gview

    @staticmethod
    def _dims_of(generator, discriminator, forward_model):
        """None at the reference dims; else the PiganDims of the three modules (BASELINE config 5: widened MLPs and
        2048-point spectra) - the engine then runs the generic-width step (include/pigan_b200.h)."""
        g_h = tuple(getattr(generator, "hidden", (512, 256)))
        d_h = tuple(getattr(discriminator, "hidden", (512, 256)))
        f_h = tuple(getattr(forward_model, "hidden", (256, 512, 1024, 512, 256)))
        S = generator.main[0].in_features
        Mt = forward_model.output_metrics_dim
        if (g_h, d_h, f_h, S, Mt) == ((512, 256), (512, 256), (256, 512, 1024, 512, 256), 250, 8):
            return None
        if discriminator.main[0].in_features != S + 4 or forward_model.output_spectrum_dim != S:
            raise ValueError("generator / discriminator / forward model disagree on the spectrum length")
        return native.make_dims(spectrum_dim=S, metrics_dim=Mt, f_hidden=f_h, g_hidden=g_h, d_hidden=d_h)

    # ------------------------------------------------------------------
    def _args(self, spectrum, params_denorm, metrics_norm, lr_g, lr_d, operand=None, center=None):
        gp, dp = self.gs.params.tensor(), self.ds.params.tensor()
        B = metrics_norm.shape[0]
        return self.engine.make_train_args(
            spectrum=native.ptr(spectrum), params_denorm=native.ptr(params_denorm),
            metrics_norm=metrics_norm.data_ptr(), spectrum_operand=native.ptr(operand),
            spectrum_center=native.ptr(center),
            batch=B, global_batch=B * self.world,
            g_params=gp.data_ptr(), g_grads=self.g_grads.data_ptr(), g_exp_avg=self.g_m.data_ptr(),
            g_exp_avg_sq=self.g_v.data_ptr(), g_bn_buffers=self.gs.bn.tensor().data_ptr(),
            g_num_batches_tracked=self.gs.nbt.tensor().data_ptr(),
            d_params=dp.data_ptr(), d_grads=self.d_grads.data_ptr(), d_exp_avg=self.d_m.data_ptr(),
            d_exp_avg_sq=self.d_v.data_ptr(),
            lr_g=float(lr_g), lr_d=float(lr_d), step=self.step_count,
            f1_idx=self.f1_idx, f2_idx=self.f2_idx, losses=self.losses.data_ptr(), **self.lam)

    @staticmethod
    def prepare_operand(spectrum: torch.Tensor, params_denorm: torch.Tensor, center: torch.Tensor) -> torch.Tensor:
        """fp16 first-layer operand [N,256] of a whole dataset (device tensors in, device tensor out): spectrum -
        center | params - 2.5 | 1 1 | 0.  Done once, like the reference's dataset normalisation
        (data_loader.py:185-219); batches are row slices of the result and cost half the bytes of the fp32 spectra."""
        for x in (spectrum, params_denorm, center):
            if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
                raise RuntimeError("prepare_operand needs contiguous fp32 CUDA tensors")
        out = torch.empty(spectrum.shape[0], 256, device=spectrum.device, dtype=torch.float16)
        native.check(native.lib.pigan_prepare_spectrum_operand(
            spectrum.data_ptr(), params_denorm.data_ptr(), center.data_ptr(), spectrum.shape[0], spectrum.shape[1],
            params_denorm.shape[1], out.data_ptr(), native.current_stream()))
        return out

    def step_prepared(self, operand: torch.Tensor, center: torch.Tensor, metrics_norm: torch.Tensor, lr_g: float,
                      lr_d: float) -> torch.Tensor:
        """step() on a batch whose operand rows were prepared by prepare_operand (same ``center`` for every batch
        and, under data parallelism, every rank)."""
        if self.wide:
            raise RuntimeError("step_prepared: the prepared fp16 operand exists at the reference widths only")
        if (not operand.is_cuda or operand.dtype != torch.float16 or not operand.is_contiguous()
                or operand.shape[1] != 256):
            raise RuntimeError("step_prepared: operand must be a contiguous fp16 CUDA tensor [B,256]")
        return self._run(None, None, metrics_norm, lr_g, lr_d, operand=operand, center=center)

    def step(self, spectrum, params_denorm, metrics_norm, lr_g: float, lr_d: float) -> torch.Tensor:
        """One D-step + G-step on device-resident fp32 tensors.  Returns the [9] device tensor of losses
        (loss_history order); no host synchronisation happens here."""
        return self._run(spectrum, params_denorm, metrics_norm, lr_g, lr_d)

    def _run(self, spectrum, params_denorm, metrics_norm, lr_g, lr_d, operand=None, center=None) -> torch.Tensor:
        for t, name in ((spectrum, "spectrum"), (params_denorm, "params_denorm"), (metrics_norm, "metrics_norm"),
                        (center, "center")):
            if t is not None and (not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError(f"NativeTrainer.step: {name} must be a contiguous fp32 CUDA tensor")
        if self.world > 1:
            self._dp_batch = check_equal_batch(metrics_norm.shape[0], getattr(self, "_dp_batch", None), self.pg)
        self.step_count += 1
        args = self._args(spectrum, params_denorm, metrics_norm, lr_g, lr_d, operand, center)
        self._pm_spectrum = spectrum
        if self.lambda_physics_metric > 0.0:
            if spectrum is None:
                raise RuntimeError("the physics-metric term needs the fp32 spectra (step(), not step_prepared())")
            args.flags = 1          # one stream: the engine is used between the phases
        if self.world == 1:
            if self.lambda_physics_metric > 0.0:
                for ph in range(7):
                    if ph == 3:
                        self._physics_metric_term(args)
                    self.engine.train_step_phase(args, ph)
                self._add_physics_metric_loss()
            else:
                self.engine.train_step(args)
            return self.losses
        if operand is None and self._center is None:
            # one centring row on all ranks (the exchanged BatchNorm sums are sums of centred pre-activations)
            c = spectrum[:512].mean(dim=0)
            dist.all_reduce(c, group=self.pg)
            self._center = c / self.world
            self.engine.set_spectrum_center(self._center)
        # data-parallel schedule: the engine's phases with NCCL all-reduces of the batch-coupled sums between
        # them (BatchNorm forward/backward statistics, gradients, loss sums) — include/pigan_b200.h
        e = self.engine
        if self.xchg is not None:
            self._step_peer(args)
            self._add_physics_metric_loss()
            return self.losses

        def run_phase(ph):
            if ph == 3 and self.lambda_physics_metric > 0.0:
                self._physics_metric_term(args)
            e.train_step_phase(args, ph)

        run_dp_step(run_phase, self._buffer, lambda t: dist.all_reduce(t, group=self.pg),
                    dp_phase_plan(e.dims.g_hidden[0], e.dims.g_hidden[1]))
        self._add_physics_metric_loss()
        return self.losses

    # ------------------------------------------------------------------ physics-metric term
    def _physics_metric_term(self, args) -> None:
        """Between phases 2 and 3: dp_extra = lambda * dL_pm/d(params_norm) for this rank's rows."""
        from . import physics as _physics
        e = self.engine
        x = self._pm_spectrum
        n = x.shape[0]
        S = e.dims.spectrum_dim
        p = e.generator_output(n)
        out = e.forward_model_forward(p)                                  # [n, S + Mt] fp32
        rec = out[:, :S].contiguous().requires_grad_(True)
        m_rec = _physics.differentiable_peak_metrics(rec)                 # [n, 4] with a backward into rec
        tgt = _physics.peak_metrics(x)
        m_tgt = torch.stack([tgt["f_res"], tgt["Q"], tgt["FoM"], tgt["S"]], dim=1)
        ok = torch.isfinite(m_rec.detach()).all(dim=1) & torch.isfinite(m_tgt).all(dim=1)
        w = torch.tensor(self.physics_metric_weights, device=x.device, dtype=torch.float32)
        diff = torch.where(ok[:, None], m_rec - m_tgt, torch.zeros_like(m_rec))
        loss = (w * diff * diff).sum() / float(n * self.world)            # mean over the GLOBAL batch
        loss.backward()
        g_out = torch.zeros_like(out)
        g_out[:, :S] = rec.grad
        dp = e.forward_model_vjp(self.fs.params.tensor(), p, g_out)
        self._pm_dp = (self.lambda_physics_metric * dp).contiguous()       # kept alive until the step has consumed it
        self._pm_loss = loss.detach()
        self.last_physics_metric_rows = ok.sum()                           # rows where all four metrics are defined
        args.dp_extra = self._pm_dp.data_ptr()

    def _add_physics_metric_loss(self) -> None:
        if self.lambda_physics_metric > 0.0:
            pm = self._pm_loss.clone()
            if self.world > 1:
                dist.all_reduce(pm, group=self.pg)
            self.last_physics_metric_loss = pm
            self.losses[1] += self.lambda_physics_metric * pm              # g_losses: the generator's total

    def _step_peer(self, args) -> None:
        """The data-parallel schedule (dp_phase_plan) with peer-memory all-reduces: this step's gradients accumulate
        in exchange slot (net, step parity); phases that consume reduced gradients get the local reduced buffer."""
        e, x, ep = self.engine, self.xchg, self.step_count
        par = ep & 1
        args.g_grads = self._slots[(0, par)].data_ptr()
        args.d_grads = self._slots[(1, par)].data_ptr()
        channel = 0
        for phase, reductions in dp_phase_plan(e.dims.g_hidden[0], e.dims.g_hidden[1]):
            if phase == 3:
                args.d_grads = self.d_grads.data_ptr()     # Adam(D) reads the reduced gradients
                if self.lambda_physics_metric > 0.0:
                    self._physics_metric_term(args)
            if phase == 6:
                args.g_grads = self.g_grads.data_ptr()
            e.train_step_phase(args, phase)
            small = [(n, sl) for n, sl in reductions if n not in ("d_grads", "g_grads")]
            todo = [(n, sl) for n, sl in reductions if n in ("d_grads", "g_grads")]
            if len(small) == 2:      # an fp32 and an fp64 buffer: one exchange for both
                todo.append(("+".join(n for n, _ in small), small))
            else:
                todo += small
            for name, sl in todo:
                ev = None
                if self.exchange_events is not None:     # profiling aid (bench.py: dp_exchange_us)
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                if name == "d_grads":
                    x.allreduce_grads(1, self.d_grads, channel, ep)
                elif name == "g_grads":
                    x.allreduce_grads(0, self.g_grads, channel, ep)
                elif isinstance(sl, list):
                    bufs = [self._buffer(n)[s_] for n, s_ in sl]
                    f32 = next(b for b in bufs if b.dtype == torch.float32)
                    f64 = next(b for b in bufs if b.dtype == torch.float64)
                    x.allreduce_small2(f32, f64, channel, ep)
                else:
                    buf = self._buffer(name)
                    x.allreduce_small(buf if sl is None else buf[sl], channel, ep)
                if ev is not None:
                    ev[1].record()
                    self.exchange_events.append((f"after_phase_{phase}:{name}", ev[0], ev[1]))
                channel += 1

    def _buffer(self, name: str) -> torch.Tensor:
        """Device buffers the data-parallel schedule reduces (dp_phase_plan)."""
        if name == "d_grads":
            return self.d_grads
        if name == "g_grads":
            return self.g_grads
        return getattr(self.engine, name)()

    # ------------------------------------------------------------------ optimiser state for checkpoints
    def export_optimizer_state(self, optimizer_g, optimizer_d) -> None:
        """Fill torch.optim.Adam state dicts (step / exp_avg / exp_avg_sq per parameter) with views of the flat
        moment buffers so optimizer.state_dict() matches what the reference checkpoints (train_pigan.py:287-294)."""
        for opt, st, m, v in ((optimizer_g, self.gs, self.g_m, self.g_v), (optimizer_d, self.ds, self.d_m, self.d_v)):
            for p, mv, vv in zip(st.params._tensors(), st.params.views_like(m), st.params.views_like(v)):
                opt.state[p] = {"step": torch.tensor(float(self.step_count)), "exp_avg": mv, "exp_avg_sq": vv}
