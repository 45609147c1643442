"""ctypes binding of liblevins_b200.so (include/levins/abi.h) -- the harness-side view of the C ABI.

This is plumbing for tests and bench.py, not the product: the product is the shared library and its C++ host
headers (le-vins_b200/host).  There is no fallback: if the library is missing or no sm_100 GPU is present,
construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .synth import INS, XYZI, XYZIT

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# LEVINS_LIB: an alternative build of the same library (A/B timing of kernel variants on one box; tests and bench only)
LIB_PATH = os.environ.get("LEVINS_LIB") or os.path.join(PKG_DIR, "liblevins_b200.so")

OK, EMPTY = 0, 1
ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_STATE = -1, -2, -3, -4
CELLS = 300 * 300

EXPORTS = [
    "levins_abi_version", "levins_config_default", "levins_create", "levins_destroy", "levins_last_error",
    "levins_status_string", "levins_set_stream", "levins_synchronize", "levins_set_extrinsics", "levins_append_scan",
    "levins_buffer_info", "levins_buffer_clear", "levins_stack_deskew", "levins_build_range_image", "levins_associate",
    "levins_keyframe", "levins_point_cloud_projection", "levins_batch_upload", "levins_batch_run",
    "levins_batch_download", "levins_batch_set_wave", "levins_set_profiling", "levins_batch_kernel_ms",
    "levins_launch_count", "levins_batch_result_pointers", "levins_result_slot_bytes", "levins_alloc_host", "levins_free_host", "levins_append_raw",
    "levins_append_wait", "levins_ins_window_index", "levins_comm_unique_id", "levins_comm_init", "levins_comm_destroy",
    "levins_gather", "levins_peer_export", "levins_peer_attach", "levins_peer_wait", "levins_peer_detach", "levins_gathered_pointer",
    "levins_gathered_download", "levins_batch_upload_packed", "levins_batch_attach_device", "levins_associate_pixels", "levins_keyframe_pixels", "levins_plane_estimation_batch",
    "levins_point_to_plane_std_batch", "levins_depth_factor_batch", "levins_chi2_culling_batch", "levins_state_interpolate", "levins_state_to_pose", "levins_pose_at",
]


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("libm_mode", C.c_int32), ("stack_interval", C.c_double),
        ("reserved_buffer_length", C.c_double), ("nearest_depth", C.c_double), ("farthest_depth", C.c_double),
        ("plane_residual_threshold", C.c_double), ("range_spread_threshold", C.c_double),
        ("nearest_distance_counts", C.c_int32), ("deskew_precision", C.c_int32), ("max_points", C.c_uint64),
        ("max_ins", C.c_uint64), ("max_features", C.c_uint64),
    ]


KF_DESC = np.dtype([("pt_begin", "<u8"), ("pt_count", "<u8"), ("ins_begin", "<u8"), ("ins_count", "<u8"),
                    ("q_begin", "<u8"), ("q_count", "<u8"), ("stamp", "<f8")])
assert KF_DESC.itemsize == 56
PACKED = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("dt_ns", "<u4")])
SCAN_DESC = np.dtype([("first", "<u8"), ("count", "<u8"), ("stamp", "<f8")])


def pack_points(points, scan_first, scan_stamp):
    """PointXYZIT records -> (packed 16-byte records, scan descriptors, the PointXYZIT records with the times the device
    will reconstruct: stamp + dt_ns * 1e-9).  scan_first: index of each scan's first point; scan_stamp: its stamp."""
    n = len(points)
    first = np.asarray(scan_first, np.uint64)
    count = np.diff(np.concatenate([first, [n]]).astype(np.uint64))
    stamp = np.asarray(scan_stamp, np.float64)
    per_point_stamp = np.repeat(stamp, count.astype(np.int64))
    dt = np.rint((points["time"] - per_point_stamp) * 1e9)
    assert dt.min() >= 0 and dt.max() < 2**32, "a point lies before its scan's stamp or more than 4.29 s after it"
    packed = np.zeros(n, PACKED)
    packed["x"], packed["y"], packed["z"], packed["dt_ns"] = points["x"], points["y"], points["z"], dt.astype(np.uint32)
    scans = np.zeros(len(first), SCAN_DESC)
    scans["first"], scans["count"], scans["stamp"] = first, count, stamp
    exact = points.copy()
    exact["time"] = per_point_stamp + packed["dt_ns"].astype(np.float64) * 1.0e-9
    exact["intensity"] = 0
    return packed, scans, exact

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
        L.levins_abi_version.restype = i32
        L.levins_config_default.argtypes = [C.POINTER(Config)]
        L.levins_create.restype = i32
        L.levins_create.argtypes = [i32, C.POINTER(Config), C.POINTER(vp)]
        L.levins_destroy.argtypes = [vp]
        L.levins_last_error.restype = C.c_char_p
        L.levins_last_error.argtypes = [vp]
        L.levins_status_string.restype = C.c_char_p
        L.levins_status_string.argtypes = [i32]
        for name, args in {
            "levins_set_stream": [vp, vp],
            "levins_synchronize": [vp],
            "levins_append_wait": [vp],
            "levins_set_extrinsics": [vp, vp, vp, dbl],
            "levins_append_scan": [vp, vp, sz],
            "levins_buffer_info": [vp, vp, vp, vp],
            "levins_buffer_clear": [vp],
            "levins_stack_deskew": [vp, dbl, vp, sz, vp, sz, vp],
            "levins_build_range_image": [vp, vp, sz, vp, sz, vp, vp],
            "levins_associate": [vp, vp, sz, vp, vp, vp, vp, vp],
            "levins_keyframe": [vp, dbl, vp, sz, vp, sz, vp, vp, vp, vp],
            "levins_associate_pixels": [vp, vp, vp, sz, i32, vp, vp, vp, vp, vp, vp],
            "levins_keyframe_pixels": [vp, dbl, vp, sz, vp, vp, sz, i32, vp, vp, vp, vp, vp, vp, vp, vp],
            "levins_point_cloud_projection": [vp, vp, vp, vp, sz],
            "levins_batch_upload": [vp, vp, sz, vp, sz, vp, sz, vp, sz],
            "levins_plane_estimation_batch": [vp, vp, sz, dbl, vp, vp, vp, vp, vp],
            "levins_point_to_plane_std_batch": [vp, vp, sz, vp, sz, dbl, vp],
            "levins_depth_factor_batch": [vp, vp, vp, sz, dbl, dbl, vp, vp, vp],
            "levins_chi2_culling_batch": [vp, vp, sz, sz, dbl, vp, vp],
            "levins_batch_attach_device": [vp, vp, sz, vp, sz, vp, sz, vp, sz],
            "levins_batch_upload_packed": [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz],
            "levins_batch_run": [vp],
            "levins_batch_download": [vp, vp, vp, vp, vp, vp],
            "levins_batch_set_wave": [vp, i32],
            "levins_batch_result_pointers": [vp, vp, vp, vp],
            "levins_set_profiling": [vp, i32],
            "levins_batch_kernel_ms": [vp, vp, vp],
            "levins_comm_unique_id": [vp],
            "levins_comm_init": [vp, vp, i32, i32, C.POINTER(vp)],
            "levins_comm_destroy": [vp],
            "levins_gather": [vp, vp, i32],
            "levins_peer_export": [vp, i32, sz, vp],
            "levins_peer_attach": [vp, i32, i32, vp],
            "levins_peer_wait": [vp],
            "levins_peer_detach": [vp],
            "levins_gathered_pointer": [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz)],
            "levins_gathered_download": [vp, vp, vp],
        }.items():
            fn = getattr(L, name)
            fn.restype = i32
            fn.argtypes = args
        L.levins_launch_count.restype = C.c_uint64
        L.levins_launch_count.argtypes = [vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_config() -> Config:
    cfg = Config()
    lib().levins_config_default(C.byref(cfg))
    return cfg


class LevinsError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"levins status {status} ({lib().levins_status_string(status).decode()}): {msg}")
        self.status = status


class Context:
    """One levins_ctx on one GPU."""

    def __init__(self, device: int = 0, cfg: Config | None = None, **overrides):
        L = lib()
        cfg = cfg or default_config()
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        h = C.c_void_p()
        st = L.levins_create(int(device), C.byref(cfg), C.byref(h))
        if st != OK:
            raise LevinsError(st, "levins_create failed (no sm_100 GPU? the library has no CPU fallback)")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().levins_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, allow_empty=True):
        if st < 0 or (st == EMPTY and not allow_empty):
            raise LevinsError(st, lib().levins_last_error(self._h).decode())
        return st

    # ---- configuration ----
    def set_stream(self, cuda_stream: int):
        self._check(lib().levins_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(lib().levins_synchronize(self._h))

    def set_extrinsics(self, pose_b_c, pose_c_l, td_b_l):
        pb = np.ascontiguousarray(pose_b_c, np.float64)
        pc = np.ascontiguousarray(pose_c_l, np.float64)
        assert pb.size == 12 and pc.size == 12
        self._check(lib().levins_set_extrinsics(self._h, _p(pb), _p(pc), float(td_b_l)))

    def set_profiling(self, on: bool):
        self._check(lib().levins_set_profiling(self._h, int(on)))

    @property
    def launch_count(self) -> int:
        return int(lib().levins_launch_count(self._h))

    # ---- streaming ----
    def append_scan(self, points):
        pts = np.ascontiguousarray(points, XYZIT)
        self._check(lib().levins_append_scan(self._h, _p(pts), len(pts)))

    def append_raw(self, fmt, records, n, timebase_ns=0, stamp=0.0, to_gps_time=True, scan_line=128, nearest=1.0, farthest=200.0,
                   want_result=True):
        """levins_append_raw: driver records (flat uint8, n * point_step bytes) -> ring.  Returns (appended, start, end); with
        want_result=False nothing is read back and the call is asynchronous (returns None)."""
        from .rawfmt import RAW_FORMAT

        f = np.ascontiguousarray(fmt, RAW_FORMAT).reshape(1)
        rec = records if (isinstance(records, np.ndarray) and records.dtype == np.uint8 and records.flags.c_contiguous) \
            else np.ascontiguousarray(records, np.uint8)
        m, a, b = C.c_size_t(0), C.c_double(0), C.c_double(0)
        L = lib()
        L.levins_append_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_double, C.c_int, C.c_int,
                                        C.c_double, C.c_double, C.POINTER(C.c_size_t), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        if not want_result:
            self._check(L.levins_append_raw(self._h, _p(f), _p(rec), int(n), int(timebase_ns), float(stamp), int(bool(to_gps_time)),
                                            int(scan_line), float(nearest), float(farthest), None, None, None))
            return None
        self._check(L.levins_append_raw(self._h, _p(f), _p(rec), int(n), int(timebase_ns), float(stamp), int(bool(to_gps_time)),
                                        int(scan_line), float(nearest), float(farthest), C.byref(m), C.byref(a), C.byref(b)))
        return m.value, a.value, b.value

    def append_wait(self):
        self._check(lib().levins_append_wait(self._h))

    def buffer_info(self):
        n, a, b = C.c_size_t(0), C.c_double(0), C.c_double(0)
        self._check(lib().levins_buffer_info(self._h, C.byref(n), C.byref(a), C.byref(b)))
        return n.value, a.value, b.value

    def buffer_clear(self):
        self._check(lib().levins_buffer_clear(self._h))

    def stack_deskew(self, stamp, ins, want_cloud=True, capacity=None):
        ins = np.ascontiguousarray(ins, INS)
        n = C.c_size_t(0)
        cap = int(capacity if capacity is not None else self.buffer_info()[0])
        out = np.zeros(max(cap, 1), XYZI) if want_cloud else None
        st = self._check(lib().levins_stack_deskew(self._h, float(stamp), _p(ins), len(ins), _p(out), cap, C.byref(n)))
        return st, n.value, (out[: n.value] if want_cloud else None)

    def build_range_image(self, cloud=None, want_unit=True, want_winner=True):
        cl = None if cloud is None else np.ascontiguousarray(cloud, XYZI)
        unit = np.zeros(CELLS, XYZI) if want_unit else None
        win = np.full(CELLS, -2, np.int32) if want_winner else None
        m = C.c_size_t(0)
        st = self._check(lib().levins_build_range_image(self._h, _p(cl), 0 if cl is None else len(cl), _p(unit), CELLS,
                                                        _p(win), C.byref(m)))
        return st, m.value, (unit[: m.value] if want_unit else None), win

    def associate(self, queries):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
        f = len(q)
        ty = np.zeros(f, np.int8)
        xd = np.zeros((f, 4), np.float32)
        ki = np.full((f, 5), -7, np.int32)
        kd = np.zeros((f, 5), np.float32)
        fl = np.zeros(f, np.int32)
        self._check(lib().levins_associate(self._h, _p(q), f, _p(ty), _p(xd), _p(ki), _p(kd), _p(fl)))
        return {"type": ty, "xyz_depth": xd, "knn_idx": ki, "knn_d2": kd, "flags": fl}

    def keyframe(self, stamp, ins, queries):
        ins = np.ascontiguousarray(ins, INS)
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
        f = len(q)
        ty = np.zeros(f, np.int8)
        xd = np.zeros((f, 4), np.float32)
        npj, m = C.c_size_t(0), C.c_size_t(0)
        st = self._check(lib().levins_keyframe(self._h, float(stamp), _p(ins), len(ins), _p(q), f, _p(ty), _p(xd),
                                               C.byref(npj), C.byref(m)))
        return {"status": st, "type": ty, "xyz_depth": xd, "n_projected": npj.value, "m": m.value}

    @staticmethod
    def _camera(intr, dist):
        cam = np.zeros(10, np.float64)
        cam[:5] = np.asarray(intr, np.float64)[:5]
        d = np.asarray(dist if dist is not None else [], np.float64)
        cam[5:5 + len(d)] = d[:5]
        return cam

    def associate_pixels(self, intr, dist, pixels, undistort=True, pose_ref=None):
        cam = self._camera(intr, dist)
        px = np.ascontiguousarray(pixels, np.float32).reshape(-1, 2)
        f = len(px)
        pose = None if pose_ref is None else np.ascontiguousarray(pose_ref, np.float64)
        out = {"type": np.zeros(f, np.int8), "xyz_depth": np.zeros((f, 4), np.float32), "accepted": np.zeros(f, np.uint8),
               "pw": np.zeros((f, 3), np.float64) if pose is not None else None, "undistorted": np.zeros((f, 2), np.float32)}
        self._check(lib().levins_associate_pixels(self._h, _p(cam), _p(px), f, int(bool(undistort)), _p(pose), _p(out["type"]),
                                                  _p(out["xyz_depth"]), _p(out["accepted"]), _p(out["pw"]), _p(out["undistorted"])))
        return out

    def keyframe_pixels(self, stamp, ins, intr, dist, pixels, undistort=True, pose_ref=None):
        ins = np.ascontiguousarray(ins, INS)
        cam = self._camera(intr, dist)
        px = np.ascontiguousarray(pixels, np.float32).reshape(-1, 2)
        f = len(px)
        pose = None if pose_ref is None else np.ascontiguousarray(pose_ref, np.float64)
        out = {"type": np.zeros(f, np.int8), "xyz_depth": np.zeros((f, 4), np.float32), "accepted": np.zeros(f, np.uint8),
               "pw": np.zeros((f, 3), np.float64) if pose is not None else None, "undistorted": np.zeros((f, 2), np.float32)}
        npj, m = C.c_size_t(0), C.c_size_t(0)
        out["status"] = self._check(lib().levins_keyframe_pixels(
            self._h, float(stamp), _p(ins), len(ins), _p(cam), _p(px), f, int(bool(undistort)), _p(pose), _p(out["type"]),
            _p(out["xyz_depth"]), _p(out["accepted"]), _p(out["pw"]), _p(out["undistorted"]), C.byref(npj), C.byref(m)))
        out["n_projected"], out["m"] = npj.value, m.value
        return out

    def plane_estimation_batch(self, points, threshold):
        pts = np.ascontiguousarray(points, XYZI).reshape(-1)
        n = len(pts) // 5
        out = {"valid": np.zeros(n, np.uint8), "unit_normal": np.zeros((n, 3)), "norm_inverse": np.zeros(n),
               "distance": np.zeros((n, 5)), "n_distance": np.zeros(n, np.int32)}
        self._check(lib().levins_plane_estimation_batch(self._h, _p(pts), n, float(threshold), _p(out["valid"]), _p(out["unit_normal"]),
                                                        _p(out["norm_inverse"]), _p(out["distance"]), _p(out["n_distance"])))
        return out

    def point_to_plane_std_batch(self, nearest, group_size, targets, point_std):
        near = np.ascontiguousarray(nearest, XYZI).reshape(-1)
        tgt = np.ascontiguousarray(targets, XYZI).reshape(-1)
        out = np.zeros(len(tgt))
        self._check(lib().levins_point_to_plane_std_batch(self._h, _p(near), int(group_size), _p(tgt), len(tgt), float(point_std), _p(out)))
        return out

    def depth_factor_batch(self, depth, invdepth, sigma=0.1, huber_a=1.0):
        d = np.ascontiguousarray(depth, np.float64)
        i = np.ascontiguousarray(invdepth, np.float64)
        out = {"residual": np.zeros(len(d)), "jacobian": np.zeros(len(d)), "cost": np.zeros(len(d))}
        self._check(lib().levins_depth_factor_batch(self._h, _p(d), _p(i), len(d), float(sigma), float(huber_a), _p(out["residual"]),
                                                    _p(out["jacobian"]), _p(out["cost"])))
        return out

    def chi2_culling_batch(self, residuals, threshold=5.991):
        r = np.ascontiguousarray(residuals, np.float64)
        n, dim = r.shape
        cost, outlier = np.zeros(n), np.zeros(n, np.uint8)
        self._check(lib().levins_chi2_culling_batch(self._h, _p(r), n, dim, float(threshold), _p(cost), _p(outlier)))
        return cost, outlier

    def point_cloud_projection(self, pose, cloud):
        pose = np.ascontiguousarray(pose, np.float64)
        src = np.ascontiguousarray(cloud, XYZI)
        dst = np.zeros(len(src), XYZI)
        self._check(lib().levins_point_cloud_projection(self._h, _p(pose), _p(src), _p(dst), len(src)))
        return dst

    # ---- batch ----
    def batch_upload(self, points, ins, queries, descs):
        self._b_points = np.ascontiguousarray(points, XYZIT)
        self._b_ins = np.ascontiguousarray(ins, INS)
        self._b_q = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
        self._b_descs = np.ascontiguousarray(descs, KF_DESC)
        self._check(lib().levins_batch_upload(self._h, _p(self._b_points), len(self._b_points), _p(self._b_ins),
                                              len(self._b_ins), _p(self._b_q), len(self._b_q), _p(self._b_descs),
                                              len(self._b_descs)))

    def batch_attach_device(self, d_points, n_points, d_ins, n_ins, d_queries, n_queries, descs):
        """device pointers (ints) of caller-owned pools; nothing is copied"""
        self._b_descs = np.ascontiguousarray(descs, KF_DESC)
        self._b_q = np.zeros((int(n_queries), 3), np.float32)  # only its length is used (batch_download)
        self._check(lib().levins_batch_attach_device(self._h, C.c_void_p(int(d_points)), int(n_points), C.c_void_p(int(d_ins)), int(n_ins),
                                                     C.c_void_p(int(d_queries)), int(n_queries), _p(self._b_descs), len(self._b_descs)))

    def batch_upload_packed(self, packed, scans, ins, queries, descs):
        self._b_points = np.ascontiguousarray(packed, PACKED)
        self._b_scans = np.ascontiguousarray(scans, SCAN_DESC)
        self._b_ins = np.ascontiguousarray(ins, INS)
        self._b_q = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
        self._b_descs = np.ascontiguousarray(descs, KF_DESC)
        self._check(lib().levins_batch_upload_packed(self._h, _p(self._b_points), len(self._b_points), _p(self._b_scans),
                                                     len(self._b_scans), _p(self._b_ins), len(self._b_ins), _p(self._b_q),
                                                     len(self._b_q), _p(self._b_descs), len(self._b_descs)))

    def batch_run(self):
        self._check(lib().levins_batch_run(self._h))

    def batch_set_wave(self, w: int):
        self._check(lib().levins_batch_set_wave(self._h, int(w)))

    def batch_download(self, out=None):
        nq, nk = len(self._b_q), len(self._b_descs)
        if out is None:
            out = {"type": np.zeros(nq, np.int8), "xyz_depth": np.zeros((nq, 4), np.float32),
                   "flags": np.zeros(nq, np.int32), "n_projected": np.zeros(nk, np.uint32), "m": np.zeros(nk, np.uint32)}
        self._check(lib().levins_batch_download(self._h, _p(out.get("type")), _p(out.get("xyz_depth")), _p(out.get("flags")),
                                                _p(out.get("n_projected")), _p(out.get("m"))))
        return out

    def batch_result_pointers(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(lib().levins_batch_result_pointers(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"type": a.value, "xyz_depth": b.value, "flags": c.value}

    # ---- exchange step (one context per rank) ----
    @staticmethod
    def comm_unique_id() -> np.ndarray:
        out = np.zeros(128, np.uint8)
        st = lib().levins_comm_unique_id(_p(out))
        if st != OK:
            raise LevinsError(st, "levins_comm_unique_id (is libnccl.so.2 loadable?)")
        return out

    def comm_init(self, unique_id, n_ranks, rank) -> int:
        uid = np.ascontiguousarray(unique_id, np.uint8)
        assert uid.size == 128
        comm = C.c_void_p()
        self._check(lib().levins_comm_init(self._h, _p(uid), int(n_ranks), int(rank), C.byref(comm)))
        return comm.value

    @staticmethod
    def comm_destroy(comm):
        lib().levins_comm_destroy(C.c_void_p(comm))

    def gather(self, comm, n_ranks):
        self._check(lib().levins_gather(self._h, C.c_void_p(comm), int(n_ranks)))

    def peer_export(self, n_ranks, n_queries_cap) -> np.ndarray:
        h = np.zeros(64, np.uint8)
        self._check(lib().levins_peer_export(self._h, int(n_ranks), int(n_queries_cap), _p(h)))
        return h

    def peer_attach(self, rank, n_ranks, all_handles):
        hs = np.ascontiguousarray(all_handles, np.uint8).reshape(-1)
        assert hs.size == 64 * n_ranks
        self._check(lib().levins_peer_attach(self._h, int(rank), int(n_ranks), _p(hs)))

    def peer_wait(self):
        self._check(lib().levins_peer_wait(self._h))

    def peer_detach(self):
        self._check(lib().levins_peer_detach(self._h))

    def gathered_download(self, n_ranks):
        p, slot, nq = C.c_void_p(), C.c_size_t(0), C.c_size_t(0)
        self._check(lib().levins_gathered_pointer(self._h, C.byref(p), C.byref(slot), C.byref(nq)))
        ty = np.zeros(n_ranks * nq.value, np.int8)
        xd = np.zeros((n_ranks * nq.value, 4), np.float32)
        self._check(lib().levins_gathered_download(self._h, _p(ty), _p(xd)))
        return {"type": ty.reshape(n_ranks, nq.value), "xyz_depth": xd.reshape(n_ranks, nq.value, 4)}

    def batch_kernel_ms(self):
        ms = np.zeros(4, np.float32)
        n = np.zeros(4, np.uint32)
        self._check(lib().levins_batch_kernel_ms(self._h, _p(ms), _p(n)))
        return ms, n


def batch_descs(b) -> np.ndarray:
    """levins_keyframe_desc array for a synth.Batch (queries laid out keyframe-major)."""
    k = b.n_keyframes
    f = b.queries.shape[1]
    d = np.zeros(k, KF_DESC)
    d["pt_begin"], d["pt_count"] = b.pt_begin, b.pt_count
    d["ins_begin"], d["ins_count"] = b.ins_begin, b.ins_count
    d["q_begin"] = np.arange(k) * f
    d["q_count"] = f
    d["stamp"] = b.stamps
    return d


class HostBuffer:
    """Page-locked host array from levins_alloc_host (scan buffers: asynchronous H2D at PCIe speed)."""

    def __init__(self, count: int, dtype=XYZIT):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(count) * self.dtype.itemsize
        p = C.c_void_p()
        L = lib()
        L.levins_alloc_host.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.levins_free_host.argtypes = [C.c_void_p]
        if L.levins_alloc_host(self.nbytes, C.byref(p)) != OK:
            raise MemoryError("levins_alloc_host failed")
        self._p = p
        self.array = np.frombuffer((C.c_char * max(self.nbytes, 1)).from_address(p.value), dtype=self.dtype, count=int(count))

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            lib().levins_free_host(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so that torch can view library-owned device memory (no copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(torch, ptr, shape, dtype, device_index):
    typestr = {torch.float32: "<f4", torch.int8: "|i1", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_CudaArray(ptr, shape, typestr), device=torch.device("cuda", device_index))
