"""The C-ABI shared library: builds, loads, exports every symbol include/levins/abi.h declares, and refuses to run
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "levins", "abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(levins_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_are_exported(built_lib):
    from levins import abi

    lib = ctypes.CDLL(built_lib)
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in abi.h but not exported"
    assert sorted(abi.EXPORTS) == names, "python binding and header disagree"
    assert lib.levins_abi_version() == 2


def test_config_defaults_are_the_reference_constants(built_lib):
    from levins import abi

    cfg = abi.default_config()
    assert cfg.struct_size == ctypes.sizeof(abi.Config)
    assert cfg.stack_interval == 0.5  # le_vins.h:130
    assert cfg.reserved_buffer_length == 10.0  # eviction at 10 + 2 * 0.5 = 11 s, le_vins.cc:156
    assert (cfg.nearest_depth, cfg.farthest_depth) == (1.0, 200.0)  # visual_mappoint.h:50-51
    assert (cfg.plane_residual_threshold, cfg.range_spread_threshold) == (0.1, 0.5)  # tracking.cc:600, 609
    assert cfg.nearest_distance_counts == 3 and cfg.libm_mode == 0


def test_record_layouts():
    from levins import abi, synth

    assert synth.XYZIT.itemsize == 32 and synth.XYZIT.fields["time"][1] == 24 and synth.XYZIT.fields["intensity"][1] == 16
    assert synth.XYZI.itemsize == 32 and synth.XYZI.fields["intensity"][1] == 16
    assert synth.INS.itemsize == 72 and abi.KF_DESC.itemsize == 56


def test_no_cpu_fallback(built_lib):
    """Without an sm_100 device the library must fail loudly; it never computes on the CPU."""
    import torch

    from levins import abi

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU box")
    with pytest.raises(abi.LevinsError) as e:
        abi.Context(0)
    assert e.value.status == abi.ERR_CUDA


def test_product_does_not_touch_the_oracle():
    """The product path (package + C ABI sources) never imports, includes or links anything under oracle/."""
    pkg = os.path.join(ROOT, "le-vins_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", ".hpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle_py" not in txt and "levins_oracle" not in txt and "liblevins_oracle" not in txt, f


def test_result_slots_are_16_byte_multiples(built_lib):
    """The stride of the per-rank slots in a gathered block: k_plane stores float4 results into every rank's slot, so the slot
    size must be a multiple of 16 for ANY query count (375 000 queries per rank once gave 6 375 000-byte slots and misaligned
    stores on the odd ranks)."""
    lib = ctypes.CDLL(built_lib)
    lib.levins_result_slot_bytes.restype = ctypes.c_size_t
    lib.levins_result_slot_bytes.argtypes = [ctypes.c_size_t]
    for n in list(range(0, 70)) + [288, 291, 6400, 375000, 375001, 10**7 + 3]:
        b = lib.levins_result_slot_bytes(n)
        assert b % 16 == 0 and b % 17 == 0 and b // 17 >= max(n, 1) and b // 17 < max(n, 1) + 16


def test_header_is_plain_c():
    """include/levins/abi.h is the drop-in boundary: it must compile as C99 (plain pointers and sizes, no C++ types) and as C++."""
    import subprocess

    hdr = os.path.join(ROOT, "include", "levins", "abi.h")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
