"""Builds libgpulin.so (hand-written CUDA for sm_100a) in-tree with nvcc.

``python -m scip_b200.build`` or ``scip_b200.build.build_library()``.  The library travels to the GPU box with the
repository snapshot; it is git-ignored.  -fmad=false: every fp64 product and sum must round exactly like the
reference's C code (no FMA contraction), see csrc/gpulin_device.cuh.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpulin.so")
SOURCES = ["gpulin.cu"]
DEPS = ["gpulin.cu", "gpulin_kernels.cuh", "gpulin_device.cuh", "gpulin_ranged.cuh", os.path.join("..", "..", "include", "gpulin.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libgpulin.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
