"""ctypes binding of the CPU oracle (oracle/liblevins_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs,
never by the product package.  See oracle/levins_oracle.h ("parity unpinned").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblevins_oracle.so")

XYZIT = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("pad0", "<f4"), ("intensity", "<f4"), ("pad1", "<f4"), ("time", "<f8")]
)
XYZI = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"), ("intensity", "<f4"), ("pad", "<f4", (3,))])
INS = np.dtype([("time", "<f8"), ("state_time", "<f8"), ("p", "<f8", (3,)), ("q", "<f8", (4,))])
CELLS = 300 * 300

LIBM_FLOAT, LIBM_DOUBLE = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "levins_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", HERE, "-B", "liblevins_oracle.so"], check=True, env=env, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
        L.orc_const_range_image_angle.restype = dbl
        L.orc_const_tan_value.restype = dbl
        L.orc_const_nearest_sq_threshold.restype = dbl
        L.orc_const_image_size.restype = i32
        L.orc_get_ins_window_index.restype = sz
        L.orc_get_ins_window_index.argtypes = [vp, sz, dbl]
        L.orc_get_pose_from_ins_window.restype = i32
        L.orc_get_pose_from_ins_window.argtypes = [vp, sz, vp, dbl, vp]
        L.orc_state_pose_interpolation.argtypes = [vp, vp, dbl, vp, vp]
        L.orc_quaternion2vector.argtypes = [vp, vp]
        L.orc_rotvec2quaternion.argtypes = [vp, vp]
        L.orc_quat_to_matrix.argtypes = [vp, vp]
        L.orc_point_projection.argtypes = [vp, vp, vp]
        L.orc_relative_projection.argtypes = [vp, vp, vp, vp]
        L.orc_relative_projection_to_camera.argtypes = [vp, vp, vp, vp, vp]
        L.orc_point_cloud_projection.argtypes = [vp, vp, vp, sz, i32]
        L.orc_point_square_distance.restype = dbl
        L.orc_point_square_distance.argtypes = [vp, vp]
        L.orc_plane_estimation.restype = i32
        L.orc_plane_estimation.argtypes = [vp, sz, dbl, vp, vp, vp, vp]
        L.orc_point_to_plane_std.restype = dbl
        L.orc_point_to_plane_std.argtypes = [vp, sz, vp, dbl]
        L.orc_ring_create.restype = vp
        L.orc_ring_destroy.argtypes = [vp]
        L.orc_ring_append.argtypes = [vp, vp, sz]
        L.orc_ring_size.restype = sz
        L.orc_ring_size.argtypes = [vp]
        L.orc_ring_data.restype = vp
        L.orc_ring_data.argtypes = [vp]
        L.orc_ring_evict_before.restype = sz
        L.orc_ring_evict_before.argtypes = [vp, dbl]
        L.orc_stack_deskew.restype = i32
        L.orc_stack_deskew.argtypes = [vp, sz, dbl, dbl, vp, sz, vp, vp, dbl, vp, vp, vp, vp, vp, i32]
        L.orc_range_image.restype = sz
        L.orc_range_image.argtypes = [vp, sz, i32, vp, vp, vp, vp]
        L.orc_kdtree_build.restype = vp
        L.orc_kdtree_build.argtypes = [vp, sz]
        L.orc_kdtree_destroy.argtypes = [vp]
        L.orc_kdtree_knn.restype = i32
        L.orc_kdtree_knn.argtypes = [vp, vp, i32, vp, vp]
        L.orc_brute_knn.restype = i32
        L.orc_brute_knn.argtypes = [vp, sz, vp, i32, vp, vp]
        L.orc_colpiv_qr_solve_5x3.argtypes = [vp, vp, vp]
        L.orc_plane_fitting.restype = i32
        L.orc_plane_fitting.argtypes = [vp, sz, vp, vp, vp, vp, vp]
        L.orc_associate.argtypes = [vp, sz, vp, vp, sz, vp, vp, vp, vp, vp]
        L.orc_pixel2cam.argtypes = [vp, C.c_float, C.c_float, vp]
        L.orc_pixel2unitcam.argtypes = [vp, C.c_float, C.c_float, vp]
        L.orc_cam2world.argtypes = [vp, vp, vp]
        L.orc_is_good_depth.restype = i32
        L.orc_is_good_depth.argtypes = [dbl, dbl]
        L.orc_atan2f_fdlibm.restype = C.c_float
        L.orc_atan2f_fdlibm.argtypes = [C.c_float, C.c_float]
        L.orc_atan2f_selfcheck.restype = C.c_long
        L.orc_atan2f_selfcheck.argtypes = [C.c_long, C.c_ulonglong]
        L.orc_keyframe.restype = sz
        L.orc_keyframe.argtypes = [vp, sz, dbl, dbl, vp, sz, vp, vp, dbl, vp, sz, i32, i32, vp, vp, vp]
        L.orc_max_threads.restype = i32
        L.orc_unix2gps_sow.restype = dbl
        L.orc_unix2gps_sow.argtypes = [dbl]
        L.orc_convert_raw.restype = sz
        L.orc_convert_raw.argtypes = [vp, vp, sz, C.c_uint64, dbl, i32, i32, dbl, dbl, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def max_threads() -> int:
    return int(lib().orc_max_threads())


def consts():
    L = lib()
    return {
        "angle": L.orc_const_range_image_angle(),
        "tan": L.orc_const_tan_value(),
        "thr2": L.orc_const_nearest_sq_threshold(),
        "size": L.orc_const_image_size(),
    }


def get_ins_window_index(ins, time) -> int:
    ins = _c(ins, INS)
    return int(lib().orc_get_ins_window_index(_p(ins), len(ins), float(time)))


def get_pose_from_ins_window(ins, pose_ext, time):
    ins = _c(ins, INS)
    ext = _c(pose_ext, np.float64)
    out = np.zeros(12)
    ok = lib().orc_get_pose_from_ins_window(_p(ins), len(ins), _p(ext), float(time), _p(out))
    return bool(ok), out


def state_pose_interpolation(s0, s1, t):
    a, b = _c(np.array([s0], dtype=INS), INS), _c(np.array([s1], dtype=INS), INS)
    p, q = np.zeros(3), np.zeros(4)
    lib().orc_state_pose_interpolation(_p(a), _p(b), float(t), _p(p), _p(q))
    return p, q


def quaternion2vector(q):
    q = _c(q, np.float64)
    r = np.zeros(3)
    lib().orc_quaternion2vector(_p(q), _p(r))
    return r


def rotvec2quaternion(r):
    r = _c(r, np.float64)
    q = np.zeros(4)
    lib().orc_rotvec2quaternion(_p(r), _p(q))
    return q


def quat_to_matrix(q):
    q = _c(q, np.float64)
    R = np.zeros(9)
    lib().orc_quat_to_matrix(_p(q), _p(R))
    return R.reshape(3, 3).T  # column-major -> numpy


def point_projection(pose, pt):
    pose = _c(pose, np.float64)
    a = _c(np.array([pt], dtype=XYZI), XYZI)
    o = np.zeros(1, dtype=XYZI)
    lib().orc_point_projection(_p(pose), _p(a), _p(o))
    return o[0]


def relative_projection(pose0, pose1, pt):
    p0, p1 = _c(pose0, np.float64), _c(pose1, np.float64)
    a = _c(np.array([pt], dtype=XYZI), XYZI)
    o = np.zeros(1, dtype=XYZI)
    lib().orc_relative_projection(_p(p0), _p(p1), _p(a), _p(o))
    return o[0]


def relative_projection_to_camera(pose_c_l, pose_cam0, pose_cam1, pt):
    pc, p0, p1 = _c(pose_c_l, np.float64), _c(pose_cam0, np.float64), _c(pose_cam1, np.float64)
    a = _c(np.array([pt], dtype=XYZI), XYZI)
    o = np.zeros(1, dtype=XYZI)
    lib().orc_relative_projection_to_camera(_p(pc), _p(p0), _p(p1), _p(a), _p(o))
    return o[0]


def point_cloud_projection(pose, cloud, nthreads=1):
    pose = _c(pose, np.float64)
    cloud = _c(cloud, XYZI)
    out = np.zeros(len(cloud), dtype=XYZI)
    lib().orc_point_cloud_projection(_p(pose), _p(cloud), _p(out), len(cloud), int(nthreads))
    return out


def plane_estimation(points, threshold):
    pts = _c(points, XYZI)
    n, ninv, dist, nd = np.zeros(3), np.zeros(1), np.zeros(5), np.zeros(1, np.int32)
    ok = lib().orc_plane_estimation(_p(pts), len(pts), float(threshold), _p(n), _p(ninv), _p(dist), _p(nd))
    return bool(ok), n, float(ninv[0]), dist[: int(nd[0])]


def point_to_plane_std(nearest, target, point_std):
    pts = _c(nearest, XYZI)
    t = _c(np.array([target], dtype=XYZI), XYZI)
    return float(lib().orc_point_to_plane_std(_p(pts), len(pts), _p(t), float(point_std)))


class Ring:
    """std::deque<PointXYZIT> lidar_point_buffer_ with the reference's append/evict rules."""

    def __init__(self):
        self._h = lib().orc_ring_create()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ring_destroy(self._h)
            self._h = None

    def append(self, pts):
        pts = _c(pts, XYZIT)
        lib().orc_ring_append(self._h, _p(pts), len(pts))

    def __len__(self):
        return int(lib().orc_ring_size(self._h))

    def evict_before(self, t):
        return int(lib().orc_ring_evict_before(self._h, float(t)))

    def data(self):
        n = len(self)
        if n == 0:
            return np.zeros(0, dtype=XYZIT)
        addr = lib().orc_ring_data(self._h)
        buf = (C.c_char * (n * XYZIT.itemsize)).from_address(addr)
        return np.frombuffer(buf, dtype=XYZIT, count=n).copy()


def stack_deskew(points, stamp, ins, pose_b_c, pose_c_l, td_b_l, stack_interval=0.5, nthreads=0, details=False):
    """le_vins.cc:392-470.  Returns dict(first, n, cloud[, ins_index, kept], ret)."""
    points = _c(points, XYZIT)
    ins = _c(ins, INS)
    pb, pc = _c(pose_b_c, np.float64), _c(pose_c_l, np.float64)
    nthreads = nthreads or max_threads()
    first, n_out = C.c_size_t(0), C.c_size_t(0)
    out = np.zeros(max(len(points), 1), dtype=XYZI)
    idx = np.full(max(len(points), 1), -2, dtype=np.int32) if details else None
    kept = np.zeros(max(len(points), 1), dtype=np.uint8) if details else None
    ret = lib().orc_stack_deskew(_p(points), len(points), float(stamp), float(stack_interval), _p(ins), len(ins), _p(pb),
                                 _p(pc), float(td_b_l), C.byref(first), C.byref(n_out), _p(out), _p(idx), _p(kept),
                                 int(nthreads))
    n = n_out.value
    res = {"first": first.value, "n": n, "cloud": out[:n], "ret": ret}
    if details:
        res["ins_index"] = idx[:n]
        res["kept"] = kept[:n]
    return res


def range_image(cloud, libm_mode=LIBM_FLOAT):
    """tracking.cc:298-358.  Returns dict(m, cell_of_point, winner, unit_sphere, cell_to_compact)."""
    cloud = _c(cloud, XYZI)
    cop = np.full(max(len(cloud), 1), -1, dtype=np.int32)
    win = np.full(CELLS, -1, dtype=np.int32)
    us = np.zeros(CELLS, dtype=XYZI)
    c2c = np.full(CELLS, -1, dtype=np.int32)
    m = lib().orc_range_image(_p(cloud), len(cloud), int(libm_mode), _p(cop), _p(win), _p(us), _p(c2c))
    return {"m": int(m), "cell_of_point": cop[: len(cloud)], "winner": win, "unit_sphere": us[:m].copy(),
            "cell_to_compact": c2c}


class KdTree:
    def __init__(self, unit_sphere):
        self.pts = _c(unit_sphere, XYZI)
        self._h = lib().orc_kdtree_build(_p(self.pts), len(self.pts))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_kdtree_destroy(self._h)
            self._h = None

    def knn(self, q, k=5):
        q = _c(q, np.float32)
        idx, d2 = np.full(k, -1, np.int32), np.full(k, np.finfo(np.float32).max, np.float32)
        n = lib().orc_kdtree_knn(self._h, _p(q), k, _p(idx), _p(d2))
        return n, idx, d2


def brute_knn(unit_sphere, q, k=5):
    pts = _c(unit_sphere, XYZI)
    q = _c(q, np.float32)
    idx, d2 = np.full(k, -1, np.int32), np.full(k, np.finfo(np.float32).max, np.float32)
    n = lib().orc_brute_knn(_p(pts), len(pts), _p(q), k, _p(idx), _p(d2))
    return n, idx, d2


def colpiv_qr_solve_5x3(A, b):
    A, b = _c(A, np.float64), _c(b, np.float64)
    x = np.zeros(3)
    lib().orc_colpiv_qr_solve_5x3(_p(A), _p(b), _p(x))
    return x


def associate(unit_sphere, queries, use_kdtree=False):
    """tracking.cc:564-634 over F queries.  Returns dict(type, xyz_depth, knn_idx, knn_d2, flags)."""
    us = _c(unit_sphere, XYZI)
    q = _c(queries, np.float32).reshape(-1, 3)
    f = len(q)
    ty = np.zeros(f, np.int8)
    xd = np.zeros((f, 4), np.float32)
    ki = np.full((f, 5), -1, np.int32)
    kd = np.zeros((f, 5), np.float32)
    fl = np.zeros(f, np.int32)
    tree = KdTree(us) if use_kdtree else None
    lib().orc_associate(_p(us), len(us), tree._h if tree else None, _p(q), f, _p(ty), _p(xd), _p(ki), _p(kd), _p(fl))
    return {"type": ty, "xyz_depth": xd, "knn_idx": ki, "knn_d2": kd, "flags": fl}


def pixel2unitcam(intr, u, v):
    intr = _c(intr, np.float64)
    o = np.zeros(3)
    lib().orc_pixel2unitcam(_p(intr), float(u), float(v), _p(o))
    return o


def pixel2cam(intr, u, v):
    intr = _c(intr, np.float64)
    o = np.zeros(3)
    lib().orc_pixel2cam(_p(intr), float(u), float(v), _p(o))
    return o


def cam2world(pc, pose):
    pc, pose = _c(pc, np.float64), _c(pose, np.float64)
    o = np.zeros(3)
    lib().orc_cam2world(_p(pc), _p(pose), _p(o))
    return o


def undistort_points(intr, dist, px):
    """Camera::undistortPoints (camera.cc:72-74) on an (F, 2) f32 pixel array; dist = k1 k2 p1 p2 k3."""
    intr = _c(intr, np.float64)
    dist = _c(np.concatenate([np.asarray(dist, np.float64), np.zeros(5)])[:5], np.float64)
    px = _c(px, np.float32).reshape(-1, 2)
    out = np.zeros_like(px)
    lib().orc_undistort_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib().orc_undistort_points(_p(intr), _p(dist), _p(px), len(px), _p(out))
    return out


def depth_factor(depth, invdepth, sigma=0.1, huber_a=1.0):
    """DepthFactor residual, Jacobian and robustified cost (depth_factor.h:38-42, optimizer.cc:659-665)."""
    L = lib()
    L.orc_depth_factor.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    r, j, c = C.c_double(0), C.c_double(0), C.c_double(0)
    L.orc_depth_factor(float(depth), float(invdepth), float(sigma), float(huber_a), C.byref(r), C.byref(j), C.byref(c))
    return r.value, j.value, c.value


def chi2_outlier(residual, threshold=5.991):
    L = lib()
    L.orc_chi2_outlier.restype = C.c_int
    L.orc_chi2_outlier.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_void_p]
    r = _c(residual, np.float64)
    c = C.c_double(0)
    out = L.orc_chi2_outlier(_p(r), len(r), float(threshold), C.byref(c))
    return bool(out), c.value


def is_good_depth(depth, scale=1.0) -> bool:
    return bool(lib().orc_is_good_depth(float(depth), float(scale)))


def atan2f_fdlibm(y, x) -> float:
    return float(lib().orc_atan2f_fdlibm(float(y), float(x)))


def atan2f_selfcheck(n, seed=0) -> int:
    return int(lib().orc_atan2f_selfcheck(int(n), int(seed)))


def keyframe(points, stamp, ins, pose_b_c, pose_c_l, td_b_l, queries, stack_interval=0.5, libm_mode=LIBM_FLOAT,
             nthreads=0):
    """Whole keyframe with the reference's structure (stage 1 parallel, stages 2-3 serial + KD-tree)."""
    points = _c(points, XYZIT)
    ins = _c(ins, INS)
    pb, pc = _c(pose_b_c, np.float64), _c(pose_c_l, np.float64)
    q = _c(queries, np.float32).reshape(-1, 3)
    f = len(q)
    ty = np.zeros(f, np.int8)
    xd = np.zeros((f, 4), np.float32)
    npj = C.c_size_t(0)
    m = lib().orc_keyframe(_p(points), len(points), float(stamp), float(stack_interval), _p(ins), len(ins), _p(pb),
                           _p(pc), float(td_b_l), _p(q), f, int(libm_mode), int(nthreads or max_threads()), _p(ty),
                           _p(xd), C.byref(npj))
    return {"m": int(m), "type": ty, "xyz_depth": xd, "n_projected": npj.value}


RAW_FORMAT = np.dtype([(k, "<u4") for k in ("kind", "point_step", "off_x", "off_y", "off_z", "off_intensity", "off_time",
                                             "off_ring", "off_tag", "time_kind", "ring_bytes", "pad")])


def unix2gps_sow(unixs):
    return float(lib().orc_unix2gps_sow(float(unixs)))


def convert_raw(fmt, records, n, timebase_ns=0, stamp=0.0, to_gps_time=True, scan_line=128, nearest=1.0, farthest=200.0):
    """ROS/lidar/lidar_converter.cc:35-256.  `fmt`: one RAW_FORMAT record; `records`: uint8 array of n * point_step bytes.
    Returns (points XYZIT, start, end)."""
    f = np.ascontiguousarray(fmt, RAW_FORMAT).reshape(1)
    rec = np.ascontiguousarray(records, np.uint8)
    out = np.zeros(max(int(n), 1), XYZIT)
    st, en = C.c_double(0), C.c_double(0)
    m = lib().orc_convert_raw(_p(f), _p(rec), int(n), int(timebase_ns), float(stamp), int(bool(to_gps_time)), int(scan_line),
                              float(nearest), float(farthest), _p(out), C.byref(st), C.byref(en))
    return out[:m], st.value, en.value
