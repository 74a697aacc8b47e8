"""Compiles the oracle's C restatement (oracle/physics.c) into oracle/_build/libpigan_oracle.so with gcc.

ORACLE = test infrastructure.  Building the checker is not using it: only tests/, smoke() and bench.py's
CPU legs load the result.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libpigan_oracle.so")
SRC = os.path.join(HERE, "physics.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # -ffp-contract=off / no fast-math: keep IEEE double evaluation order identical to NumPy's
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"gcc failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build())
