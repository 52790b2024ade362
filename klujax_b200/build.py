"""Builds the CUDA shared library of the package in-tree (sm_100a only).

    python -m klujax_b200.build [--force]

Output: klujax_b200/_lib/libklujax_cuda.so (git-ignored, travels with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OUT = os.path.join(OUT_DIR, "libklujax_cuda.so")
SOURCES = ["klujax_cuda.cu", "ordering.cpp", "seed_lu.cpp", "plan.cpp", "rowreg.cpp", "rowreg_ptx.cpp",
           "klujax_cuda_ffi.cc"]  # the FFI shim compiles to a stub where the XLA headers are absent
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))) + \
          [os.path.join("..", "..", "include", "klujax_cuda.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "-shared", "--expt-relaxed-constexpr"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not (force or _stale()):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [NVCC] + FLAGS + os.environ.get("KLUJAX_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
