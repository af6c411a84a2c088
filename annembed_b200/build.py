"""In-tree nvcc build of libannembed_cuda.so for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libannembed_cuda.so")
SOURCES = ["annembed_cuda.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))) + [os.path.join("..", "..", "include", "annembed_cuda.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]
# Development builds: `-split-compile 0` halves the build time (1.5 instead of 3.5 minutes), but the code ptxas generates
# for k_sweep_events was then seen to vary from build to build of the SAME source (96 registers with or without 160 bytes of
# spills: 27 % difference in time per embed, profiles/r02_ab_build_flags.txt).  The shipped library is built without it.
FAST_BUILD_FLAGS = ["-split-compile", "0"]


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
    # `verbose` prints the resource usage of every kernel afterwards (cuobjdump), not ptxas -v
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    if verbose:
        cuobjdump = os.path.join(os.path.dirname(_nvcc()), "cuobjdump")
        subprocess.call([cuobjdump, "--dump-resource-usage", target])
    return target


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
