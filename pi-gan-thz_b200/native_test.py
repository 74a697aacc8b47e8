"""ctypes binding of libpigan_b200_test.so: the test hooks of the tcgen05 GEMM core (include/pigan_b200_debug.h).

Separate from the product library on purpose - only tests/ and tools/ import this module."""
from __future__ import annotations

import ctypes as C
import os

from .native import PIGAN_OK, PiganError, current_stream, ptr  # noqa: F401  (re-exported for the tests)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpigan_b200_test.so")

if not os.path.exists(LIB_PATH):
    import importlib.util

    _spec = importlib.util.spec_from_file_location("_pigan_build", os.path.join(_HERE, "build.py"))
    _mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
    _mod.build()
lib = C.CDLL(LIB_PATH)

_vp, _i32 = C.c_void_p, C.c_int32
SIGNATURES = {
    "pigan_debug_force_streamed": (_i32, [_i32]),
    "pigan_debug_gemm_tn": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "pigan_debug_linear": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "pigan_debug_linear2": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "pigan_debug_gemm_nt": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype, _fn.argtypes = _res, _args
lib.pigan_last_error.restype = C.c_char_p


def check(code: int) -> None:
    if code != PIGAN_OK:
        raise PiganError(code, (lib.pigan_last_error() or b"").decode())
