"""ctypes binding of libpigan_b200.so — the only way host code reaches the CUDA kernels.

There is no fallback: if the library is missing (and cannot be built because nvcc is absent) importing
this module raises, and every compute entry point fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpigan_b200.so")

PIGAN_OK = 0


class PiganError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pigan_b200 error {code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        import importlib.util

        spec = importlib.util.spec_from_file_location("_pigan_build", os.path.join(_HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return C.CDLL(LIB_PATH)


lib = _load()

_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f32 = C.c_float


class PiganDims(C.Structure):
    _fields_ = [
        ("spectrum_dim", _i32),
        ("param_dim", _i32),
        ("metrics_dim", _i32),
        ("g_hidden", _i32 * 2),
        ("d_hidden", _i32 * 2),
        ("f_hidden", _i32 * 5),
    ]


class PiganFwdTrainArgs(C.Structure):
    """include/pigan_b200.h: struct PiganFwdTrainArgs (surrogate training step)."""
    _fields_ = [
        ("params_norm", C.c_void_p), ("spectrum", C.c_void_p), ("metrics_norm", C.c_void_p),
        ("batch", C.c_int64), ("global_batch", C.c_int64), ("first_row", C.c_int64),
        ("f_params", C.c_void_p), ("f_grads", C.c_void_p), ("f_exp_avg", C.c_void_p), ("f_exp_avg_sq", C.c_void_p),
        ("lr", C.c_float), ("step", C.c_int64),
        ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("max_norm", C.c_float),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64),
        ("losses", C.c_void_p), ("loss_sums", C.c_void_p), ("mask_dump", C.c_void_p),
    ]


class PiganTrainArgs(C.Structure):
    _fields_ = [
        ("spectrum", _vp), ("params_denorm", _vp), ("metrics_norm", _vp),
        ("batch", _i64), ("global_batch", _i64),
        ("g_params", _vp), ("g_grads", _vp), ("g_exp_avg", _vp), ("g_exp_avg_sq", _vp),
        ("g_bn_buffers", _vp), ("g_num_batches_tracked", _vp),
        ("d_params", _vp), ("d_grads", _vp), ("d_exp_avg", _vp), ("d_exp_avg_sq", _vp),
        ("lr_g", _f32), ("lr_d", _f32), ("step", _i64),
        ("lambda_recon", _f32), ("lambda_physics_spectrum", _f32), ("lambda_physics_metrics", _f32),
        ("lambda_maxwell", _f32), ("lambda_lc", _f32), ("lambda_param_range", _f32), ("lambda_bnn_kl", _f32),
        ("f1_idx", _i32), ("f2_idx", _i32),
        ("losses", _vp),
        ("spectrum_operand", _vp), ("spectrum_center", _vp),
        ("dp_extra", _vp), ("flags", _i32),
    ]


# name -> (restype, argtypes); every symbol include/pigan_b200.h declares must appear here
SIGNATURES = {
    "pigan_abi_version": (_i32, []),
    "pigan_last_error": (C.c_char_p, []),
    "pigan_launch_count": (_i64, []),
    "pigan_dp_region_bytes": (C.c_size_t, [_i64]),
    "pigan_dp_grad_slot_offset": (C.c_size_t, [_i64, _i32, _i32]),
    "pigan_dp_alloc": (_i32, [C.c_size_t, C.POINTER(_vp)]),
    "pigan_dp_free": (_i32, [_vp]),
    "pigan_dp_ipc_export": (_i32, [_vp, _vp]),
    "pigan_dp_ipc_open": (_i32, [_vp, C.POINTER(_vp)]),
    "pigan_dp_ipc_close": (_i32, [_vp]),
    "pigan_dp_create": (_i32, [C.POINTER(_vp), _i32, _i32, C.POINTER(_vp), _i64]),
    "pigan_dp_destroy": (_i32, [_vp]),
    "pigan_dp_allreduce_small": (_i32, [_vp, _vp, _i32, _i32, _i32, C.c_uint32, _vp]),
    "pigan_dp_allreduce_small2": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, C.c_uint32, _vp]),
    "pigan_dp_allreduce_grads": (_i32, [_vp, _i32, _vp, _i64, _i32, C.c_uint32, _vp, _vp]),
    "pigan_engine_profile_begin": (_i32, [_vp, C.c_char_p]),
    "pigan_engine_profile_end": (_i32, [_vp, C.c_char_p, C.c_size_t]),
    "pigan_default_dims": (None, [C.POINTER(PiganDims)]),
    "pigan_generator_param_count": (_i64, [C.POINTER(PiganDims)]),
    "pigan_discriminator_param_count": (_i64, [C.POINTER(PiganDims)]),
    "pigan_forward_model_param_count": (_i64, [C.POINTER(PiganDims)]),
    "pigan_generator_bn_buffer_count": (_i64, [C.POINTER(PiganDims)]),
    "pigan_engine_workspace_bytes": (C.c_size_t, [C.POINTER(PiganDims), _i64]),
    "pigan_engine_create": (_i32, [C.POINTER(_vp), C.POINTER(PiganDims), _i64, _vp, C.c_size_t, _vp]),
    "pigan_engine_destroy": (_i32, [_vp]),
    "pigan_engine_load_forward_model": (_i32, [_vp, _vp, _vp]),
    "pigan_engine_set_spectrum_center": (_i32, [_vp, _vp]),
    "pigan_prepare_spectrum_operand": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "pigan_generator_forward": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "pigan_discriminator_forward": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "pigan_generator_backward": (_i32, [_vp, _vp, _vp, _i64, _vp, _f32, _vp, _vp]),
    "pigan_discriminator_backward": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _f32, _vp, _vp, _vp]),
    "pigan_forward_model_forward": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "pigan_train_step": (_i32, [_vp, C.POINTER(PiganTrainArgs), _vp]),
    "pigan_train_step_phase": (_i32, [_vp, C.POINTER(PiganTrainArgs), _i32, _vp]),
    "pigan_engine_generator_output": (_vp, [_vp]),
    "pigan_engine_bn_sums": (_vp, [_vp]),
    "pigan_engine_bn_bwd_sums": (_vp, [_vp]),
    "pigan_engine_loss_sums": (_vp, [_vp]),
    "pigan_score_candidates": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "pigan_validate_model": (_i32, [_vp, _vp, _vp, _vp, _vp, _f32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "pigan_search_workspace_bytes": (C.c_size_t, [_vp, _i32]),
    "pigan_inverse_design_search": (_i32, [_vp, _vp, _vp, _vp, _f32, C.c_uint64, _i64, _i64, _i32, _vp, _vp, _vp, _vp,
                                           _vp, C.c_size_t, _vp]),
    "pigan_fwd_train_workspace_bytes": (C.c_size_t, [_vp]),
    "pigan_forward_model_input_grad": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp, C.c_size_t, _vp]),
    "pigan_forward_model_vjp": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, C.c_size_t, _vp]),
    "pigan_fwd_train_step": (_i32, [_vp, C.POINTER(PiganFwdTrainArgs), _vp, C.c_size_t, _vp]),
    "pigan_fwd_train_step_phase": (_i32, [_vp, C.POINTER(PiganFwdTrainArgs), _i32, _vp, C.c_size_t, _vp]),
    "pigan_eval_workspace_bytes": (C.c_size_t, [_i32]),
    "pigan_regression_sums": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, C.c_size_t, _vp]),
    "pigan_regression_finalize": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "pigan_score_summary_sums": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _vp, C.c_size_t, _vp]),
    "pigan_score_summary_finalize": (_i32, [_vp, _i64, _vp, _vp]),
    "pigan_gather_rows": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp]),
    "pigan_generate_spectra": (_i32, [_vp, _vp, _vp, _i64, _i32, _f32, C.c_uint64, _i64, _i32, _vp, _vp, _vp]),
    "pigan_topk_workspace_bytes": (C.c_size_t, [_i64, _i32]),
    "pigan_topk_smallest": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, C.c_size_t, _vp]),
    "pigan_physics_metrics": (_i32, [_vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _vp]),
    "pigan_engine_trace_layernorm": (_i32, [_vp]),
    "pigan_physics_metrics_backward": (_i32, [_vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp]),
}


def _bind() -> None:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


_bind()


def last_error() -> str:
    return (lib.pigan_last_error() or b"").decode()


def check(code: int) -> None:
    if code != PIGAN_OK:
        raise PiganError(code, last_error())


def ptr(t) -> int:
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def default_dims() -> PiganDims:
    d = PiganDims()
    lib.pigan_default_dims(C.byref(d))
    return d


def make_dims(spectrum_dim=None, metrics_dim=None, f_hidden=None, g_hidden=None, d_hidden=None) -> PiganDims:
    """The reference dims with some widths replaced (BASELINE config 5: hidden 2048, 2048-point spectra).  Engines
    with non-reference dims serve the surrogate's entry points only (include/pigan_b200.h)."""
    d = default_dims()
    if spectrum_dim is not None:
        d.spectrum_dim = int(spectrum_dim)
    if metrics_dim is not None:
        d.metrics_dim = int(metrics_dim)
    for name, val, n in (("f_hidden", f_hidden, 5), ("g_hidden", g_hidden, 2), ("d_hidden", d_hidden, 2)):
        if val is not None:
            if len(val) != n:
                raise ValueError(f"{name} needs {n} widths")
            setattr(d, name, (_i32 * n)(*[int(v) for v in val]))
    return d


def dims_key(d: PiganDims) -> tuple:
    return (d.spectrum_dim, d.param_dim, d.metrics_dim, tuple(d.g_hidden), tuple(d.d_hidden), tuple(d.f_hidden))
