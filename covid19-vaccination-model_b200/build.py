"""Builds libvaxmc.so (hand-written sm_100a CUDA + the C ABI of include/vaxmc.h) in-tree.

`nvcc` cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with
the gpurun snapshot.  There is exactly one target architecture: sm_100a.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "vaxmc.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "vaxmc_kernels.cuh"), os.path.join(ROOT, "include", "vaxmc.h")]
LIB = os.environ.get("VAXMC_LIB") or os.path.join(HERE, "libvaxmc.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvaxmc.so cannot be built (and there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("VAXMC_NVCC_EXTRA", "").split()
    # build next to the target and rename: ranks of one torchrun job that all find the library
    # stale may compile concurrently, and none of them may ever load a half-written file
    tmp = f"{LIB}.tmp{os.getpid()}"
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra, "-o", tmp, SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB)
    if verbose:
        print(proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
