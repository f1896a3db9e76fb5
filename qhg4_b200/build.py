"""Build the CUDA C-ABI library in-tree: qhg4_b200/libqhg_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libqhg_b200.so")
SOURCES = ["qhg_pop.cu"]
HEADERS = ["qhg_kernels.cuh", "qhg_cells.cuh", "qhg_decide.cuh", "qhg_genes.cuh", "qhg_rng.cuh", os.path.join("..", "..", "include", "qhg_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v", "-ccbin", "/usr/bin/g++", "-I/usr/include", "-ldl"]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    extra = os.environ.get("QHG_NVCC_EXTRA", "").split()
    cmd = [NVCC] + FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
