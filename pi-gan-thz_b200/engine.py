"""Python host wrapper of the C-ABI engine (include/pigan_b200.h).  PyTorch is plumbing here: it owns the
device memory and the stream; every computation goes through ``native.lib``."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import native
from .native import PiganTrainArgs, check, lib

LOSS_KEYS = ["d_losses", "g_losses", "adv_losses", "recon_spec_losses", "recon_metrics_losses", "maxwell_losses",
             "lc_losses", "param_range_losses", "bnn_kl_losses"]


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"pigan_b200: {what} must live on a CUDA device — this path has no CPU fallback")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Engine:
    """One engine per (process, device).  ``max_batch`` bounds the rows of any later call."""

    def __init__(self, max_batch: int, device: Optional[torch.device] = None, dims: Optional[native.PiganDims] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("pigan_b200: no CUDA device — the B200 path has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.max_batch = int(max_batch)
        self.dims = dims if dims is not None else native.default_dims()
        with torch.cuda.device(self.device):
            nbytes = lib.pigan_engine_workspace_bytes(C.byref(self.dims), self.max_batch)
            if nbytes == 0:
                raise native.PiganError(-3, "unsupported dimensions")
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            handle = C.c_void_p()
            check(lib.pigan_engine_create(C.byref(handle), C.byref(self.dims), self.max_batch,
                                          self.workspace.data_ptr(), nbytes, native.current_stream()))
        self.handle = handle
        self._f_params = None
        self._topk_ws = None
        self._search_ws = None

    def __del__(self):
        h = getattr(self, "handle", None)
        if h and lib is not None:  # lib is None while the interpreter shuts down
            lib.pigan_engine_destroy(h)
            self.handle = None

    # ------------------------------------------------------------------ weights
    def load_forward_model(self, f_flat: torch.Tensor) -> None:
        _require_cuda(f_flat, "forward-model parameters")
        self._f_params = f_flat  # keep alive: the engine reads biases / norm weights from it
        check(lib.pigan_engine_load_forward_model(self.handle, f_flat.data_ptr(), native.current_stream()))

    def set_spectrum_center(self, center: Optional[torch.Tensor]) -> None:
        """One centring row for every rank of a data-parallel job (see include/pigan_b200.h); None = per call."""
        if center is not None:
            _require_cuda(center, "center")
            center = _f32c(center).reshape(-1)
        self._center = center  # keep alive
        check(lib.pigan_engine_set_spectrum_center(self.handle, native.ptr(center)))

    # ------------------------------------------------------------------ module forwards
    def generator_forward(self, g_flat, bn, nbt, spectrum, training: bool) -> torch.Tensor:
        _require_cuda(spectrum, "spectrum")
        x = _f32c(spectrum)
        out = torch.empty(x.shape[0], self.dims.param_dim, device=x.device, dtype=torch.float32)
        check(lib.pigan_generator_forward(self.handle, g_flat.data_ptr(), bn.data_ptr(), nbt.data_ptr(), x.data_ptr(),
                                          x.shape[0], int(training), out.data_ptr(), native.current_stream()))
        return out

    def discriminator_forward(self, d_flat, spectrum, params) -> torch.Tensor:
        _require_cuda(spectrum, "spectrum")
        x, p = _f32c(spectrum), _f32c(params)
        out = torch.empty(x.shape[0], 1, device=x.device, dtype=torch.float32)
        check(lib.pigan_discriminator_forward(self.handle, d_flat.data_ptr(), x.data_ptr(), p.data_ptr(), x.shape[0],
                                              out.data_ptr(), native.current_stream()))
        return out

    @staticmethod
    def _grad_scale(g: torch.Tensor) -> float:
        """s with s * max|g| = 1 (the fp16 gradient tensors of the backward kernels hold s x the gradient); one host
        read - these entry points serve autograd callers, not the fused train step."""
        m = float(g.abs().max())
        return 1.0 / m if m > 0.0 and m == m and m != float("inf") else 1.0

    def generator_backward(self, g_flat, spectrum, grad_params_norm) -> torch.Tensor:
        """Flat gradient (layout of g_flat) of sum(grad_params_norm * G(spectrum)), train-mode BatchNorm."""
        _require_cuda(spectrum, "spectrum")
        x, g = _f32c(spectrum), _f32c(grad_params_norm)
        out = torch.empty_like(g_flat)
        check(lib.pigan_generator_backward(self.handle, g_flat.data_ptr(), x.data_ptr(), x.shape[0], g.data_ptr(),
                                           self._grad_scale(g), out.data_ptr(), native.current_stream()))
        return out

    def discriminator_backward(self, d_flat, spectrum, params, grad_prob, want_grad_params: bool = True):
        """(flat gradient of D's parameters, gradient with respect to `params` [n,4] or None) of
        sum(grad_prob * D(spectrum, params))."""
        _require_cuda(spectrum, "spectrum")
        x, p, g = _f32c(spectrum), _f32c(params), _f32c(grad_prob).reshape(-1)
        out = torch.empty_like(d_flat)
        gp = torch.empty_like(p) if want_grad_params else None
        check(lib.pigan_discriminator_backward(self.handle, d_flat.data_ptr(), x.data_ptr(), p.data_ptr(), x.shape[0],
                                               g.data_ptr(), self._grad_scale(g), out.data_ptr(), native.ptr(gp),
                                               native.current_stream()))
        return out, gp

    def forward_model_forward(self, params_norm) -> torch.Tensor:
        _require_cuda(params_norm, "params_norm")
        p = _f32c(params_norm)
        out = torch.empty(p.shape[0], self.dims.spectrum_dim + self.dims.metrics_dim, device=p.device,
                          dtype=torch.float32)
        check(lib.pigan_forward_model_forward(self.handle, p.data_ptr(), p.shape[0], out.data_ptr(),
                                              native.current_stream()))
        return out

    def forward_model_input_grad(self, f_flat, params_norm, spectrum, metrics_norm, w_spectrum: float = 1.0,
                                 w_metrics: float = 0.0):
        """(dL/d params_norm [n,4], losses [2]) for L = w_spectrum * MSE(F(p).spectrum, spectrum) + w_metrics *
        MSE(F(p).metrics, metrics_norm) with the surrogate's weights frozen and Dropout off — the physics-loss gradient
        of the reference's UnifiedTrainer (unified_trainer.py:240-256, 325)."""
        _require_cuda(params_norm, "params_norm")
        p, sp, mn = _f32c(params_norm), _f32c(spectrum), _f32c(metrics_norm)
        n = p.shape[0]
        if getattr(self, "_ftrain_ws", None) is None:
            self._ftrain_ws = torch.empty(lib.pigan_fwd_train_workspace_bytes(self.handle), dtype=torch.uint8,
                                          device=self.device)
        dp = torch.empty(n, self.dims.param_dim, device=p.device, dtype=torch.float32)
        losses = torch.empty(2, device=p.device, dtype=torch.float32)
        check(lib.pigan_forward_model_input_grad(self.handle, f_flat.data_ptr(), p.data_ptr(), sp.data_ptr(),
                                                 mn.data_ptr(), n, float(w_spectrum), float(w_metrics), dp.data_ptr(),
                                                 losses.data_ptr(), self._ftrain_ws.data_ptr(),
                                                 self._ftrain_ws.numel(), native.current_stream()))
        return dp, losses

    def forward_model_vjp(self, f_flat, params_norm, grad_out) -> torch.Tensor:
        """dL/d params_norm [n,4] from dL/d F(params_norm) [n, S+Mt] (spectrum columns, then metrics), weights frozen,
        Dropout off - how a loss on F(G(x)) reaches the generator (unified_trainer.py:240-256, 325)."""
        _require_cuda(params_norm, "params_norm")
        p, g = _f32c(params_norm), _f32c(grad_out)
        n = p.shape[0]
        if g.shape != (n, self.dims.spectrum_dim + self.dims.metrics_dim):
            raise ValueError("grad_out must be [n, spectrum_dim + metrics_dim]")
        if getattr(self, "_ftrain_ws", None) is None:
            self._ftrain_ws = torch.empty(lib.pigan_fwd_train_workspace_bytes(self.handle), dtype=torch.uint8,
                                          device=self.device)
        dp = torch.empty(n, self.dims.param_dim, device=p.device, dtype=torch.float32)
        check(lib.pigan_forward_model_vjp(self.handle, f_flat.data_ptr(), p.data_ptr(), g.data_ptr(), n, dp.data_ptr(),
                                          self._ftrain_ws.data_ptr(), self._ftrain_ws.numel(),
                                          native.current_stream()))
        return dp

    def generator_output(self, n: int) -> torch.Tensor:
        """[n, P] view of the generator output params_norm of the current train step (valid after phase 2)."""
        return self._wrap(lib.pigan_engine_generator_output(self.handle), n * self.dims.param_dim,
                          torch.float32).view(n, self.dims.param_dim)

    # ------------------------------------------------------------------ training
    def make_train_args(self, **kw) -> PiganTrainArgs:
        a = PiganTrainArgs()
        for k, v in kw.items():
            setattr(a, k, v)
        return a

    def train_step(self, args: PiganTrainArgs) -> None:
        check(lib.pigan_train_step(self.handle, C.byref(args), native.current_stream()))

    def train_step_phase(self, args: PiganTrainArgs, phase: int) -> None:
        check(lib.pigan_train_step_phase(self.handle, C.byref(args), phase, native.current_stream()))

    def _wrap(self, ptr: int, n: int, dtype) -> torch.Tensor:
        """Tensor view of an engine-owned scratch region (lives inside self.workspace)."""
        off = ptr - self.workspace.data_ptr()
        es = torch.empty((), dtype=dtype).element_size()
        return self.workspace[off:off + n * es].view(dtype)

    def bn_sums(self) -> torch.Tensor:
        n = 2 * (self.dims.g_hidden[0] + self.dims.g_hidden[1])
        return self._wrap(lib.pigan_engine_bn_sums(self.handle), n, torch.float32)

    def bn_bwd_sums(self) -> torch.Tensor:
        n = 2 * (self.dims.g_hidden[0] + self.dims.g_hidden[1])
        return self._wrap(lib.pigan_engine_bn_bwd_sums(self.handle), n, torch.float32)

    def loss_sums(self) -> torch.Tensor:
        return self._wrap(lib.pigan_engine_loss_sums(self.handle), 16, torch.float64)

    # ------------------------------------------------------------------ instrumentation
    def profile_begin(self, sections=None) -> None:
        csv = ",".join(sections).encode() if sections else None
        check(lib.pigan_engine_profile_begin(self.handle, csv))

    def profile_end(self) -> Dict[str, tuple]:
        """{section: (count, total_ms)} measured with CUDA events on the launching stream."""
        buf = C.create_string_buffer(16384)
        check(lib.pigan_engine_profile_end(self.handle, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    # ------------------------------------------------------------------ scoring
    def score_candidates(self, g_flat, bn, spectra=None, target=None, noise=None, sigma: float = 0.01,
                         want_params=True) -> Dict[str, torch.Tensor]:
        ref = spectra if spectra is not None else noise
        _require_cuda(ref, "candidate spectra / noise")
        n = ref.shape[0]
        dev = ref.device
        out_p = torch.empty(n, self.dims.param_dim, device=dev, dtype=torch.float32) if want_params else None
        viol = torch.empty(n, device=dev, dtype=torch.int32)
        err = torch.empty(n, device=dev, dtype=torch.float32)
        cons = torch.empty(n, device=dev, dtype=torch.float32)
        sp = _f32c(spectra) if spectra is not None else None
        tg = _f32c(target).reshape(-1) if target is not None else None
        nz = _f32c(noise) if noise is not None else None
        check(lib.pigan_score_candidates(self.handle, g_flat.data_ptr(), bn.data_ptr(), native.ptr(sp), native.ptr(tg),
                                         native.ptr(nz), float(sigma), n, native.ptr(out_p), viol.data_ptr(),
                                         err.data_ptr(), cons.data_ptr(), native.current_stream()))
        return {"params_norm": out_p, "violations": viol, "recon_error": err, "consistency": cons}


def _validate(self, g_flat, bn, spectra, noise, sigma: float = 0.01) -> Dict[str, torch.Tensor]:
    """Loop body of UnifiedEvaluator.evaluate_model_validation (unified_evaluator.py:439-468) for all rows at once:
    per-row cycle-consistency error, prediction stability under ``sigma * noise`` and plausibility score."""
    _require_cuda(spectra, "spectra")
    n, dev = spectra.shape[0], spectra.device
    out_p = torch.empty(n, self.dims.param_dim, device=dev, dtype=torch.float32)
    cyc, stab, plaus = (torch.empty(n, device=dev, dtype=torch.float32) for _ in range(3))
    sp, nz = _f32c(spectra), _f32c(noise)
    check(lib.pigan_validate_model(self.handle, g_flat.data_ptr(), bn.data_ptr(), sp.data_ptr(), nz.data_ptr(),
                                   float(sigma), n, out_p.data_ptr(), cyc.data_ptr(), stab.data_ptr(),
                                   plaus.data_ptr(), native.current_stream()))
    return {"params_norm": out_p, "cycle_error": cyc, "stability": stab, "plausibility": plaus}


Engine.validate = _validate


def _search(self, g_flat, bn, target, sigma, seed, first, count, k, dump_noise=False):
    """pigan_inverse_design_search on this engine: (scores [k], indices [k], params_norm [k,4][, noise])."""
    dev = self.device
    nbytes = lib.pigan_search_workspace_bytes(self.handle, k)
    if nbytes == 0:
        raise native.PiganError(-1, f"k={k} out of range (1..4096)")
    if self._search_ws is None or self._search_ws.numel() < nbytes:
        self._search_ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out_s = torch.empty(k, device=dev, dtype=torch.float32)
    out_i = torch.empty(k, device=dev, dtype=torch.int64)
    out_p = torch.empty(k, 4, device=dev, dtype=torch.float32)
    noise = torch.empty(count, self.dims.spectrum_dim, device=dev, dtype=torch.float32) if dump_noise else None
    tg = _f32c(target).reshape(-1)
    check(lib.pigan_inverse_design_search(self.handle, g_flat.data_ptr(), bn.data_ptr(), tg.data_ptr(), float(sigma),
                                          int(seed), int(first), int(count), int(k), out_s.data_ptr(),
                                          out_i.data_ptr(), out_p.data_ptr(), native.ptr(noise),
                                          self._search_ws.data_ptr(), self._search_ws.numel(),
                                          native.current_stream()))
    return (out_s, out_i, out_p, noise) if dump_noise else (out_s, out_i, out_p)


Engine.search = _search


def topk_smallest(scores: torch.Tensor, k: int, index_base: int = 0, in_indices: Optional[torch.Tensor] = None):
    """k smallest scores (ascending) and their indices, on the device (csrc/topk.cu)."""
    _require_cuda(scores, "scores")
    s = _f32c(scores).reshape(-1)
    n = s.numel()
    nbytes = lib.pigan_topk_workspace_bytes(n, k)
    if nbytes == 0:
        raise native.PiganError(-1, f"k={k} out of range (1..4096)")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=s.device)
    out_s = torch.empty(k, device=s.device, dtype=torch.float32)
    out_i = torch.empty(k, device=s.device, dtype=torch.int64)
    idx = in_indices.contiguous() if in_indices is not None else None
    check(lib.pigan_topk_smallest(s.data_ptr(), native.ptr(idx), n, k, int(index_base), out_s.data_ptr(),
                                  out_i.data_ptr(), ws.data_ptr(), nbytes, native.current_stream()))
    return out_s, out_i


def launch_count() -> int:
    """CUDA kernels launched by the native library in this process so far."""
    return int(lib.pigan_launch_count())


_ENGINES: Dict[tuple, Engine] = {}


def get_engine(device, min_batch: int, dims: Optional[native.PiganDims] = None) -> Engine:
    """Process-wide engine per (device, dims), grown (re-created) when a larger batch shows up."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, native.dims_key(dims) if dims is not None else None)
    eng = _ENGINES.get(key)
    if eng is None or eng.max_batch < min_batch:
        cap = max(int(min_batch), 1024)
        if eng is not None:
            cap = max(cap, 2 * eng.max_batch)
        eng = Engine(cap, torch.device("cuda", idx), dims)
        _ENGINES[key] = eng
    return eng
