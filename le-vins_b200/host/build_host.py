"""Build the C++ host-mirror test driver (g++, links liblevins_b200.so in-tree)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(HERE, "test_host")
OUT_CPU = os.path.join(HERE, "test_host_cpu")


def build_cpu(force=False):
    """The header-only part of the mirror (no CUDA at run time); still links the library because device.h references it."""
    src = os.path.join(HERE, "test_host_cpu.cc")
    deps = [src] + [os.path.join(b, f) for b, _, fs in os.walk(HERE) for f in fs if f.endswith(".h")]
    if not force and os.path.exists(OUT_CPU) and all(os.path.getmtime(d) < os.path.getmtime(OUT_CPU) for d in deps):
        return OUT_CPU
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-I", HERE, "-I", os.path.join(ROOT, "include"), src, "-o",
           OUT_CPU, "-L", PKG, "-llevins_b200", "-Wl,-rpath," + PKG, "-pthread"]
    env = dict(os.environ)
    env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed: " + " ".join(cmd))
    return OUT_CPU


def build(force=False):
    build_cpu(force)
    src = os.path.join(HERE, "test_host.cc")
    deps = [src] + [os.path.join(b, f) for b, _, fs in os.walk(HERE) for f in fs if f.endswith(".h")]
    lib = os.path.join(PKG, "liblevins_b200.so")
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) < os.path.getmtime(OUT) for d in deps + [lib]):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", HERE, "-I", os.path.join(ROOT, "include"), src, "-o", OUT,
           "-L", PKG, "-llevins_b200", "-Wl,-rpath," + PKG, "-Wl,-rpath,$ORIGIN/..", "-pthread"]
    env = dict(os.environ)
    env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed: " + " ".join(cmd))
    return OUT


if __name__ == "__main__":
    print(build(force=True))
