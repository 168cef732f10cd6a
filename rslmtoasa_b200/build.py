"""In-tree build of the CUDA library (sm_100a only; nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "librsrec.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; librsrec.so cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    hdr = os.path.join(os.path.dirname(_HERE), "include", "rsrec.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr])


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "rsrec.cu")]
        env = dict(os.environ)
        # the image exports CC/CXX pointing at a wrapper without libgomp; nvcc must use the system g++
        subprocess.check_call(cmd + ["-ccbin", shutil.which("g++") or "/usr/bin/g++"], env=env)
    return LIB
