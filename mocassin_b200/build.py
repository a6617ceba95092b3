"""Build libmocassin_b200.so (sm_100a) in-tree with nvcc.

`python -m mocassin_b200.build` or `build()`; nvcc cross-compiles without a GPU.
Flags that matter for correctness: -fmad=false (no FMA contraction: every float32
branch decision must match the CPU oracle bit for bit), default -prec-div/-prec-sqrt.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmocassin_b200.so")
SOURCES = ["capi.cu", "transport.cu", "wavefront.cu", "tables.cu", "dust.cu"]
HEADERS = ["types.h", "detmath.cuh", "philox.cuh", "transport_core.cuh", "wf_rec.cuh", "locate.cuh", "dust.h", os.path.join("..", "..", "include", "mcb200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread",
    "-shared",
    "-ldl",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build_variant(out: str, defines=()) -> str:
    """A/B build of the same sources with extra -D flags into another file (loaded through
    $MCB200_LIB, mocassin_b200/_lib.py); the product library is `build()`."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", os.path.abspath(out)] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmocassin_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:          # python -m mocassin_b200.build --variant out.so MCB_FLY_OCC=5 ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
