"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/pigan_b200.h declares.
No compute call is made here (no GPU in the CPU suite)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="pigan_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pigan_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from pigan_b200 import native
    syms = _declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(native.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in native.SIGNATURES, f"{s} has no ctypes signature in native.py"
    for s in native.SIGNATURES:
        assert s in syms, f"{s} bound in native.py but not declared in the header"


def test_test_hooks_live_in_their_own_library():
    """include/pigan_b200_debug.h is bound by native_test.py against libpigan_b200_test.so; the product library
    exports none of the pigan_debug_* symbols."""
    from pigan_b200 import native, native_test
    syms = [s for s in _declared_symbols("pigan_b200_debug.h") if s.startswith("pigan_debug_")]
    assert len(syms) >= 4
    product = ctypes.CDLL(native.LIB_PATH)
    hooks = ctypes.CDLL(native_test.LIB_PATH)
    for s in syms:
        assert hasattr(hooks, s) and s in native_test.SIGNATURES, s
        assert not hasattr(product, s), f"{s} leaked into the product library"
    assert not any(s.startswith("pigan_debug_") for s in _declared_symbols())


def test_version_and_layout_counts():
    from pigan_b200 import native
    assert native.lib.pigan_abi_version() == 9
    d = native.default_dims()
    assert (d.spectrum_dim, d.param_dim, d.metrics_dim) == (250, 4, 8)
    # parameter counts of the reference modules
    assert native.lib.pigan_generator_param_count(None) == 262404
    assert native.lib.pigan_discriminator_param_count(None) == 262145
    assert native.lib.pigan_forward_model_param_count(None) == 1385730
    assert native.lib.pigan_generator_bn_buffer_count(None) == 2 * 512 + 2 * 256
    assert native.lib.pigan_engine_workspace_bytes(None, 65536) > 0


def test_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without CUDA the compute entry points return an error, they do not compute."""
    import torch
    if torch.cuda.is_available():
        return
    from pigan_b200 import native
    buf = (ctypes.c_float * 1024)()
    rc = native.lib.pigan_physics_metrics(ctypes.addressof(buf), 1, 250, ctypes.addressof(buf), None, 0.0, None,
                                          ctypes.addressof(buf), None)
    assert rc != 0 and native.last_error()
    handle = ctypes.c_void_p()
    rc = native.lib.pigan_engine_create(ctypes.byref(handle), None, 128, ctypes.addressof(buf), 4096, None)
    assert rc != 0


def test_new_entry_points_validate_arguments_and_refuse_to_run_without_a_gpu():
    """Data pipeline, evaluator reductions, surrogate training: bad arguments -> PIGAN_ERR_INVALID with a message;
    on a box without CUDA the kernels are not launched (error, not a CPU computation)."""
    import torch
    from pigan_b200 import native
    lib = native.lib
    buf = (ctypes.c_float * 4096)()
    p = ctypes.addressof(buf)
    assert lib.pigan_eval_workspace_bytes(250) >= 148 * 4 * 250 * 8 * 8 and lib.pigan_eval_workspace_bytes(0) == 0
    assert lib.pigan_regression_sums(None, p, 4, 4, p, 0, p, 1 << 20, None) != 0 and native.last_error()
    assert lib.pigan_regression_sums(p, p, 0, 4, p, 0, p, 1 << 20, None) != 0
    assert lib.pigan_regression_sums(p, p, 4, 4, p, 0, p, 16, None) != 0          # workspace too small
    assert lib.pigan_regression_finalize(None, 4, 4, p, None) != 0
    assert lib.pigan_score_summary_finalize(p, 0, p, None) != 0
    assert lib.pigan_gather_rows(None, 4, 16, p, 1, p, None, None) != 0
    assert lib.pigan_gather_rows(p, 4, 0, p, 1, p, None, None) != 0
    assert lib.pigan_generate_spectra(None, p, None, 4, 250, 0.1, 1, 0, 1, p, None, None) != 0
    assert lib.pigan_generate_spectra(None, p, p, 4, 4096, 0.1, 1, 0, 1, p, None, None) != 0    # S > 512
    assert lib.pigan_fwd_train_workspace_bytes(None) == 0
    assert lib.pigan_fwd_train_step(None, None, p, 1 << 20, None) != 0
    if not torch.cuda.is_available():
        assert lib.pigan_regression_sums(p, p, 4, 4, p, 0, p, 1 << 26, None) != 0
        assert lib.pigan_gather_rows(p, 4, 16, p, 1, p, None, None) != 0
        assert lib.pigan_generate_spectra(None, p, p, 4, 250, 0.1, 1, 0, 1, p, None, None) != 0
        assert "no CUDA device" in native.last_error()


def test_sass_contains_blackwell_instructions():
    """The shipped .so really is tcgen05 + TMA code (B200_PROFILING.md 'what proves a Blackwell-native kernel')."""
    import shutil
    import subprocess
    from pigan_b200 import native
    if shutil.which("cuobjdump") is None:
        return
    sass = subprocess.run(["cuobjdump", "-sass", native.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass
