"""Build liblevins_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "levins_abi.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", f) for f in ("kernels.cuh", "device_math.cuh", "exchange.cuh", "exchange_host.inc")] + [
    os.path.join(ROOT, "include", "levins", "abi.h")]
OUT = os.path.join(HERE, "liblevins_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # integer / index / validity results must reproduce the reference's non-contracted SSE2 arithmetic;
    # fused multiply-adds are requested explicitly (fma()) where only a tolerance applies
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-ldl",
]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
        env = dict(os.environ)
        env.pop("CC", None)  # the image exports a gcc wrapper as CC; nvcc should pick the PATH host compiler
        env.pop("CXX", None)
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
