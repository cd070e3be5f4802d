"""Build the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtone_b200.so")
SOURCES = ["engine.cu", "server.cu"]
HEADERS = ["common.cuh", "gemm_tc.cuh", "gemm_ref.cuh", "kernels.cuh", "ctc_phrase.cuh", "state_io.cuh", "ff_fused.cuh", "att_fused.cuh", "rowgemm.cuh",
           os.path.join("..", "..", "include", "tone_b200.h")]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_prof() -> str:
    """Diagnostic variant with in-kernel timestamps (-DTONE_PROF); never used by the product path or the tests."""
    lib = os.path.join(HERE, "libtone_b200_prof.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-DTONE_PROF",
           "-Xcompiler", "-fPIC", "-shared", "-o", lib] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    if "--prof" in sys.argv:
        print(build_prof())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
