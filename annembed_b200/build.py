"""In-tree nvcc build of libannembed_cuda.so for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libannembed_cuda.so")
SOURCES = ["annembed_cuda.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "annembed_cuda.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-split-compile", "0",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libannembed_cuda.so")
    return nvcc


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), out: str | None = None) -> str:
    """Build the library; `extra_flags`/`out` build an experiment variant beside it (e.g. -DANNEMBED_MINB_D2=4)."""
    target = out or LIB
    if not force and out is None and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
