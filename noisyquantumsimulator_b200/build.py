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
DEPS = ["rydcz.cu", "rydcz_core.cuh", "rydcz_closed.cuh", "rydcz_cluster.cuh", "rydcz_repeat.cuh", "rydcz_prologue.cuh", "rydcz_systems_d3.inc", "rydcz_systems_d4.inc",
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
    # one builder at a time (pytest-xdist workers, ranks of one node), and the library appears atomically: a process
    # that loads it while another one builds never sees a half-written file
    import fcntl
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():
            return LIB
        tmp = LIB + f".tmp{os.getpid()}"
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", tmp] + SOURCES
        if verbose:
            print("[rydcz build]", " ".join(cmd), flush=True)
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(os.path.join(CSRC, "build.log"), "w") as f:
            f.write(log)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            sys.stderr.write(log)
            raise RuntimeError("nvcc failed building librydcz.so")
        os.replace(tmp, LIB)
    return LIB


# ---- the PyTorch face of the C ABI: torch.ops.rydcz.* (csrc/rydcz_torch.cpp) -------------------------------
TORCH_LIB = os.path.join(CSRC, "librydcz_torch.so")
TORCH_DEPS = ["rydcz_torch.cpp", os.path.join("..", "..", "include", "rydcz.h")]


def torch_is_stale() -> bool:
    if not os.path.exists(TORCH_LIB):
        return True
    t = os.path.getmtime(TORCH_LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in TORCH_DEPS)


def build_torch_extension(force: bool = False, verbose: bool = True) -> str:
    """Compile csrc/rydcz_torch.cpp -> csrc/librydcz_torch.so (host C++ only: a dispatcher registration that
    calls the C ABI of librydcz.so next to it; the kernels are all in librydcz.so)."""
    if not force and not torch_is_stale():
        return TORCH_LIB
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (["g++", "-O2", "-std=c++17", "-shared", "-fPIC",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
           + [f"-I{p}" for p in inc] + ["rydcz_torch.cpp", "-o", TORCH_LIB, f"-L{tlib}", f"-L{CSRC}",
                                        "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-lrydcz",
                                        "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"])
    if verbose:
        print("[rydcz build]", " ".join(cmd), flush=True)
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building librydcz_torch.so")
    return TORCH_LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    build_torch_extension(force="--force" in sys.argv)
    print(LIB)
    print(TORCH_LIB)
