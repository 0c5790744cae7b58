"""Build librydcz.so (sm_100a) in-tree with nvcc.  No Python needed at run time beyond ctypes."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "librydcz.so")
SOURCES = ["rydcz.cu"]
DEPS = ["rydcz.cu", "rydcz_core.cuh", "rydcz_cluster.cuh", "rydcz_repeat.cuh", "rydcz_prologue.cuh", "rydcz_systems_d3.inc", "rydcz_systems_d4.inc",
        "rydcz_systems_d5.inc",
        os.path.join("..", "..", "include", "rydcz.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC,-fopenmp", "-Xptxas", "-v", "-lgomp"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build librydcz.so")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile csrc/rydcz.cu -> csrc/librydcz.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + SOURCES
    if verbose:
        print("[rydcz build]", " ".join(cmd), flush=True)
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building librydcz.so")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
