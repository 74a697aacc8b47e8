"""Builds lib/libpigan_b200.so (sm_100a only) from csrc/*.cu with nvcc.

In-tree on purpose: the .so travels to the GPU box with the repo snapshot; nothing is JIT-compiled there.
Usage: python pi-gan-thz_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpigan_b200.so")
TEST_LIB = os.path.join(LIBDIR, "libpigan_b200_test.so")   # GEMM test hooks (include/pigan_b200_debug.h): tests/ and tools/ only
TEST_ONLY = {"debug_gemm.cu"}                              # sources that stay out of the product library
TEST_SHARED = {"host_util.cu"}                             # ... and what the test library needs besides them
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    # function-local statics of templates (launch_gemm's "attribute set" flags) must stay per library: as GNU-unique
    # symbols they would be shared process-wide between the product and the test library, which both instantiate
    # some of the same kernels - one library's flag would then skip the other's cudaFuncSetAttribute
    "-Xcompiler", "-fno-gnu-unique",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return exe


def _deps_mtime() -> float:
    newest = 0.0
    for root in (CSRC, INCLUDE):
        for name in os.listdir(root):
            if name.endswith((".h", ".cuh")):
                newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def _compile(src: str, obj: str, verbose: bool) -> None:
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdr_time = _deps_mtime()
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time)
        if stale:
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = [ex.submit(_compile, s, o, verbose) for s, o in jobs]
            for f in futs:
                f.result()
    def obj_of(name):
        return os.path.join(OBJ, name[:-3] + ".o")
    for lib, members in ((LIB, [obj_of(x) for x in srcs if x not in TEST_ONLY]),
                         (TEST_LIB, [obj_of(x) for x in srcs if x in TEST_ONLY or x in TEST_SHARED])):
        need_link = bool(jobs) or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in members)
        if need_link:
            # -Bsymbolic: both libraries instantiate some of the same kernel templates; each must bind (and register
            # with the CUDA runtime) its OWN host stubs, not whichever copy the dynamic loader saw first
            cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xlinker", "-Bsymbolic",
                   "-o", lib, *members]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
