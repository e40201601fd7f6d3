"""Build the CUDA shared library in-tree (liana_py_b200/_lib/liblrperm.so) for sm_100a.

    python -m liana_py_b200.build [--force]

nvcc cross-compiles without a GPU; the built .so travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SOURCES = [os.path.join(_PKG, "csrc", "lrperm_api.cu")]
HEADERS = [os.path.join(_PKG, "csrc", "lrperm_kernels.cuh"), os.path.join(_PKG, "csrc", "lrperm_trimean.cuh"),
           os.path.join(_ROOT, "include", "lrperm.h"),
           os.path.join(_ROOT, "include", "lrperm_philox.h")]
OUT = os.path.join(_PKG, "_lib", "liblrperm.so")
OUT_DEBUG = os.path.join(_PKG, "_lib", "liblrperm_debug.so")      # same ABI + the kernels' own bounds checks (--debug)

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def is_stale(out: str = OUT) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    out = OUT_DEBUG if debug else OUT
    if not force and not is_stale(out):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [find_nvcc(), *NVCC_FLAGS, *(["-DLRPERM_BOUNDS_CHECK"] if debug else []),
           "-I", os.path.join(_ROOT, "include"), "-I", os.path.join(_PKG, "csrc"), *SOURCES, "-o", out]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
