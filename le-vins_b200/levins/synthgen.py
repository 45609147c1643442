"""HARNESS, not product: the BASELINE config-5 workload (many independent Ouster keyframes generated on the device from
per-keyframe seeds) and its numpy twin.

`DeviceGenerator` drives le-vins_b200/harness/synthgen.cu (libsynthgen_b200.so); `keyframe_numpy` computes the same keyframe
on the host with the same integer hashing and the same IEEE operations in the same order, so the two agree bit for bit
(tests/test_synthgen.py) and the CPU oracle can be run on exactly what the GPU processed.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .synth import INS, XYZIT, rig_fusionportable

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SRC = os.path.join(PKG, "harness", "synthgen.cu")
LIB = os.path.join(PKG, "libsynthgen_b200.so")

RINGS, COLS = 128, 1024
PER_SCAN = RINGS * COLS
SCAN_PERIOD, INS_DT = 0.1, 0.005
T_BASE = 345600.0


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                        "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC], check=True, env=env, capture_output=True)
    return LIB


def directions() -> np.ndarray:
    """[1024 columns][128 rings][3] unit directions of an OS1-128 (+-22.5 deg), column-major in time, f32."""
    el = np.deg2rad(np.linspace(-22.5, 22.5, RINGS))
    az = -2 * np.pi * (np.arange(COLS) + 0.5) / COLS + np.pi
    azg, elg = np.meshgrid(az, el, indexing="ij")
    d = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], axis=-1).reshape(-1, 3)
    return np.ascontiguousarray(d, np.float32)


def _mix(h):
    h = np.asarray(h, np.uint32).copy()
    with np.errstate(over="ignore"):
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x7FEB352D)
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x846CA68B)
        h ^= h >> np.uint32(16)
    return h


def _hash3(a, b, c):
    return _mix(_mix(_mix(np.uint32(a)) ^ np.asarray(b, np.uint32)) ^ np.asarray(c, np.uint32))


def _u01(h):
    return (np.asarray(h, np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


class Spec:
    """One config-5 workload: Ouster keyframes with the reference's default 0.5 s window (5 scans) or the 20-scan variant."""

    def __init__(self, seed=20261024, scans_per_kf=5, n_feat=300, stack_interval=None):
        self.seed = int(seed)
        self.scans_per_kf = int(scans_per_kf)
        self.n_feat = int(n_feat)
        self.stack_interval = float(stack_interval if stack_interval is not None else scans_per_kf * SCAN_PERIOD)
        self.rig = rig_fusionportable()
        self.td_b_l = float(self.rig.td_b_l)
        # the INS window: 0.6 s before the first point's INS time to 0.1 s past the stamp
        self.n_ins = int(round((0.6 + self.scans_per_kf * SCAN_PERIOD + 0.1 + max(0.0, -self.td_b_l) + 0.1) / INS_DT)) + 1
        self.kf_period = (self.scans_per_kf + 1) * SCAN_PERIOD  # windows do not overlap
        self.per_kf = self.scans_per_kf * PER_SCAN

    def stamp(self, gk):
        """keyframe stamp: just before the end of its last scan, so that the whole buffer slice falls into the window"""
        t0k = T_BASE + np.float64(gk) * self.kf_period
        return float(t0k + self.scans_per_kf * SCAN_PERIOD - 0.0005)


def keyframe_numpy(spec: Spec, gk: int, dirs=None):
    """(points XYZIT[per_kf], ins INS[n_ins], queries f32[n_feat, 3]) of global keyframe gk, bit-identical to the device's."""
    dirs = directions() if dirs is None else dirs
    f32 = np.float32
    seed = np.uint32(spec.seed)
    i = np.arange(spec.per_kf, dtype=np.uint32)
    scan = i // np.uint32(PER_SCAN)
    in_scan = i - scan * np.uint32(PER_SCAN)
    col = in_scan // np.uint32(RINGS)
    half_w = f32(4.0) + f32(8.0) * _u01(_hash3(seed, gk, 1))
    front = f32(15.0) + f32(60.0) * _u01(_hash3(seed, gk, 2))
    back = f32(10.0) + f32(40.0) * _u01(_hash3(seed, gk, 3))
    ground = f32(1.5) + f32(0.8) * _u01(_hash3(seed, gk, 4))
    roof = f32(4.0) + f32(20.0) * _u01(_hash3(seed, gk, 5))
    d = dirs[in_scan]
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    r = np.full(spec.per_kf, 250.0, f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(dz < f32(-1e-3), np.minimum(r, ground / -dz), r)
        r = np.where(dz > f32(1e-3), np.minimum(r, roof / dz), r)
        r = np.where(dy > f32(1e-3), np.minimum(r, half_w / dy), r)
        r = np.where(dy < f32(-1e-3), np.minimum(r, half_w / -dy), r)
        r = np.where(dx > f32(1e-3), np.minimum(r, front / dx), r)
        r = np.where(dx < f32(-1e-3), np.minimum(r, back / -dx), r)
    r = r.astype(f32)
    h = _hash3(seed ^ np.uint32(0x9E3779B9), gk, i)
    r = (r + (_u01(h) * f32(0.02) - f32(0.01))).astype(f32)
    r = np.where((h & np.uint32(1023)) == 7, f32(0.3) + f32(0.5) * _u01(_mix(h)), r).astype(f32)
    pts = np.zeros(spec.per_kf, XYZIT)
    pts["x"], pts["y"], pts["z"] = dx * r, dy * r, dz * r
    pts["intensity"] = (h & np.uint32(255)).astype(f32)
    t0k = T_BASE + np.float64(gk) * spec.kf_period
    pts["time"] = (t0k + scan.astype(np.float64) * SCAN_PERIOD) + col.astype(np.float64) * (SCAN_PERIOD / float(COLS))
    # INS
    k = np.arange(spec.n_ins, dtype=np.float64)
    t = ((t0k + spec.td_b_l) - 0.6) + k * INS_DT
    tau = t - t0k
    vx = 2.0 + 12.0 * np.float64(_u01(_hash3(seed, gk, 6)))
    vy = -0.5 + 1.0 * np.float64(_u01(_hash3(seed, gk, 7)))
    w = -0.15 + 0.3 * np.float64(_u01(_hash3(seed, gk, 8)))
    u = w * tau
    n = np.sqrt(1.0 + u * u)
    ins = np.zeros(spec.n_ins, INS)
    ins["time"] = t
    ins["state_time"] = t
    ins["p"][:, 0], ins["p"][:, 1], ins["p"][:, 2] = vx * tau, vy * tau, 0.02 * tau
    ins["q"][:, 2], ins["q"][:, 3] = u / n, 1.0 / n
    # queries
    j = np.arange(spec.n_feat, dtype=np.uint32)
    s2 = seed ^ np.uint32(0x51ED270B)
    hx, hy = _hash3(s2, gk, np.uint32(2) * j), _hash3(s2, gk, np.uint32(2) * j + np.uint32(1))
    wdt, hgt = spec.rig.resolution
    px = f32(5.0) + (hx % np.uint32(wdt - 10)).astype(f32)
    py = f32(5.0) + (hy % np.uint32(hgt - 10)).astype(f32)
    fx, fy, cx, cy = [float(v) for v in spec.rig.intr[:4]]
    y = (py.astype(np.float64) - cy) / fy
    x = (px.astype(np.float64) - cx) / fx
    s = np.sqrt((x * x + y * y) + 1.0)
    q = np.stack([x / s, y / s, 1.0 / s], axis=-1).astype(f32)
    return pts, ins, q


class DeviceGenerator:
    """Generates a run of keyframes into torch CUDA tensors (device memory the caller then hands to levins_batch_attach_device)."""

    def __init__(self, torch, device_index, spec: Spec):
        build()
        self.torch, self.dev, self.spec = torch, device_index, spec
        self.lib = C.CDLL(LIB)
        self.lib.synthgen_ouster_batch.restype = C.c_int
        self.lib.synthgen_ouster_batch.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                   C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.POINTER(C.c_double), C.c_uint32, C.c_uint32]
        self.dirs = torch.from_numpy(directions()).to(torch.device("cuda", device_index))

    def generate(self, first_kf, n_kf):
        """-> (points uint8[n_kf * per_kf * 32], ins uint8[n_kf * n_ins * 72], queries f32[n_kf * n_feat * 3]) on the device"""
        torch, sp = self.torch, self.spec
        dev = torch.device("cuda", self.dev)
        pts = torch.empty(n_kf * sp.per_kf * 32, dtype=torch.uint8, device=dev)
        ins = torch.empty(n_kf * sp.n_ins * 72, dtype=torch.uint8, device=dev)
        q = torch.empty(n_kf * sp.n_feat * 3, dtype=torch.float32, device=dev)
        intr = (C.c_double * 4)(*[float(v) for v in sp.rig.intr[:4]])
        wdt, hgt = sp.rig.resolution
        rc = self.lib.synthgen_ouster_batch(self.dev, C.c_void_p(torch.cuda.current_stream().cuda_stream), sp.seed, int(first_kf), int(n_kf),
                                            sp.scans_per_kf, sp.n_ins, sp.n_feat, T_BASE, sp.kf_period, sp.td_b_l,
                                            C.c_void_p(self.dirs.data_ptr()), C.c_void_p(pts.data_ptr()), C.c_void_p(ins.data_ptr()),
                                            C.c_void_p(q.data_ptr()), intr, int(wdt), int(hgt))
        if rc:
            raise RuntimeError(f"synthgen_ouster_batch: cudaError {rc}")
        return pts, ins, q

    def descs(self, first_kf, n_kf):
        from .abi import KF_DESC

        sp = self.spec
        d = np.zeros(n_kf, KF_DESC)
        k = np.arange(n_kf)
        d["pt_begin"], d["pt_count"] = k * sp.per_kf, sp.per_kf
        d["ins_begin"], d["ins_count"] = k * sp.n_ins, sp.n_ins
        d["q_begin"], d["q_count"] = k * sp.n_feat, sp.n_feat
        d["stamp"] = [sp.stamp(first_kf + int(i)) for i in k]
        return d
