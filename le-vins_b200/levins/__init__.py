"""Harness-side Python helpers for liblevins_b200 (ctypes binding + synthetic generators).

The product is the CUDA shared library (../liblevins_b200.so, built from ../csrc) and the C++ host headers
(../host).  Nothing in here computes results on the CPU.
"""
from . import synth  # noqa: F401
