"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/pigan_b200.h declares.
No compute call is made here (no GPU in the CPU suite)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pigan_b200.h")).read()
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


def test_version_and_layout_counts():
    from pigan_b200 import native
    assert native.lib.pigan_abi_version() == 5
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
