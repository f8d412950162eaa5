"""Build libb200comp.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m image_transformation_b200.build [--force] [--verbose]

The shared library lands in image_transformation_b200/_lib/ (git-ignored, but it
travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIBDIR, "libb200comp.so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

CUDA_SOURCES = ["b200comp.cu", "host_api.cu"]
CXX_SOURCES = ["coeffs.cpp"]
HEADERS = sorted(os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h"))) + [
    os.path.join(INCLUDE, "b200comp.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]
# the coefficient tables must round like scalar C: no fast-math, no contraction
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libb200comp.so cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in CUDA_SOURCES + CXX_SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    for s in CXX_SOURCES:
        o = os.path.join(objdir, s + ".o")
        run([os.environ.get("CXX", "g++")] + CXX_FLAGS + ["-c", os.path.join(CSRC, s), "-o", o])
        objs.append(o)
    for s in CUDA_SOURCES:
        o = os.path.join(objdir, s + ".o")
        extra = ["-Xptxas", "-v"] if verbose else []
        run([nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, s), "-o", o])
        objs.append(o)
    tmp = LIB + ".tmp"
    run([nvcc, "-shared", "-o", tmp] + objs + ["-cudart", "static", "-lpthread"])
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
