"""Build libfgk_b200.so (sm_100a) in-tree with nvcc.

    python -m flow_guided_krylov_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = ["fgk_ham.cu", "fgk_index.cu", "fgk_projh.cu", "fgk_projh4.cu", "fgk_spmv.cu", "fgk_pt2.cu", "fgk_peer.cu"]
HEADERS = ["fgk_core.cuh", "fgk_internal.cuh", "fgk_lists.cuh", "fgk_tables.h", os.path.join("..", "..", "include", "fgk_b200.h")]
LIB = os.path.join(CSRC, "libfgk_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
