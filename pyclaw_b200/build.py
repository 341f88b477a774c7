"""Builds pyclaw_b200/csrc/libclawb200.so in-tree with nvcc for sm_100a.

    python -m pyclaw_b200.build [--force]

-fmad=false is part of the numerical contract (bit-for-bit agreement with the
reference's Fortran, which is compiled without FMA contraction); -lineinfo keeps the
ncu source page usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libclawb200.so")
SOURCES = ["clawb200.cu"]
HEADERS = ["arith.cuh", "rp.cuh", "classic.cuh", "sharpclaw.cuh", os.path.join("..", "..", "include", "clawb200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
