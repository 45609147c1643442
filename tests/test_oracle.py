"""CPU tests of the oracle (no GPU): the restatement is pinned against independent implementations
(scipy Slerp / cKDTree, cv2.flann KDTREE_SINGLE, numpy lstsq, a numpy range image) and against geometry that is
known by construction.  The reference itself ships no tests or golden vectors for this path ("parity unpinned").
"""
import collections

import numpy as np
import pytest
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation, Slerp

from levins import synth


# ----------------------------------------------------------------------------------------------
# constants and libm
# ----------------------------------------------------------------------------------------------
def test_constants_match_reference_expressions(oracle):
    c = oracle.consts()
    assert c["size"] == 300  # tracking.h:126
    assert c["angle"] == 0.3 * np.pi / 180.0
    assert c["tan"].hex() == "0x1.fffffffffffffp-1"  # tan(pi/4) in double
    assert c["thr2"] == 2.4673785519540605e-4  # (3 sin 0.3deg)^2, tracking.h:130-133


def test_atan2f_restatement_equals_libm(oracle):
    """The fdlibm-lineage atan2f that both the oracle documents and the GPU kernel re-implements is bit-identical to
    this image's glibc atan2f (what the reference's float-overload calls would execute)."""
    assert oracle.atan2f_selfcheck(20_000_000, 12345) == 0
    import ctypes
    import ctypes.util

    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.atan2f.restype = ctypes.c_float
    libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
    for y, x in [(0.0, 1.0), (1.0, 0.0), (-1.0, 0.0), (0.0, -1.0), (-0.0, -1.0), (1e-30, 1.0), (3.0, -4.0), (1.0, 1.0),
                 (2.5, 1e-9), (float("inf"), 1.0), (1.0, float("inf")), (0.4375, 1.0), (2.4375, 1.0)]:
        assert oracle.atan2f_fdlibm(y, x) == libm.atan2f(y, x)


# ----------------------------------------------------------------------------------------------
# INS window lookup and interpolation (misc.cc:25-103, rotation.cc:59-68)
# ----------------------------------------------------------------------------------------------
def _ins(n=120, rate=200.0, seed=0):
    rng = np.random.default_rng(seed)
    ins = np.zeros(n, synth.INS)
    ins["time"] = 1000.0 + np.arange(n) / rate
    ins["state_time"] = ins["time"]
    ins["p"] = np.cumsum(rng.normal(0, 0.01, (n, 3)), axis=0) + [10, -5, 2]
    r = Rotation.from_rotvec(np.cumsum(rng.normal(0, 0.004, (n, 3)), axis=0))
    ins["q"] = r.as_quat()
    return ins


def test_ins_window_index_semantics(oracle):
    ins = _ins()
    t = ins["time"]
    # inside: t[i-1] <= time < t[i]
    for time in [t[0], t[0] + 1e-9, t[5], (t[5] + t[6]) / 2, np.nextafter(t[6], 0), t[-2], np.nextafter(t[-1], 0)]:
        i = oracle.get_ins_window_index(ins, time)
        assert i == np.searchsorted(t, time, side="right")
        assert t[i - 1] <= time < t[i]
    # outside or on the newest sample: 0 (misc.cc:28-30)
    for time in [t[0] - 1e-6, t[-1], t[-1] + 1.0]:
        assert oracle.get_ins_window_index(ins, time) == 0
    assert oracle.get_ins_window_index(ins[:1], t[0]) == 0
    assert oracle.get_ins_window_index(ins[:0], t[0]) == 0


def test_rotvec_quaternion_roundtrip(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        rv = rng.normal(0, 1.0, 3) * rng.choice([1e-9, 1e-3, 0.5, 2.0])
        q = oracle.rotvec2quaternion(rv)
        assert np.allclose(q, Rotation.from_rotvec(rv).as_quat(), atol=1e-15)
        back = oracle.quaternion2vector(q)
        # Eigen's AngleAxis(quaternion) returns the angle in [0, pi] (2 atan2(|v|, |w|)): same rotation, maybe other branch
        assert np.allclose(Rotation.from_rotvec(back).as_matrix(), Rotation.from_rotvec(rv).as_matrix(), atol=1e-12)
        assert np.linalg.norm(back) <= np.pi + 1e-12
        if np.linalg.norm(rv) < np.pi:
            assert np.allclose(back, rv, rtol=1e-9, atol=1e-15)
    assert np.array_equal(oracle.quaternion2vector([0, 0, 0, 1.0]), [0, 0, 0])
    qn = np.array([0.1, -0.2, 0.3, 0.9])
    qn /= np.linalg.norm(qn)
    assert np.allclose(oracle.quat_to_matrix(qn), Rotation.from_quat(qn).as_matrix(), atol=1e-15)


def test_pose_interpolation_vs_slerp(oracle):
    ins = _ins(seed=2)
    rig = synth.rig_r3live()
    rng = np.random.default_rng(3)
    for _ in range(100):
        i = int(rng.integers(1, len(ins)))
        s = rng.random()
        t = ins["time"][i - 1] + s * (ins["time"][i] - ins["time"][i - 1])
        ok, pose = oracle.get_pose_from_ins_window(ins, rig.pose_b_c, t)
        assert ok
        R = pose[:9].reshape(3, 3).T
        sl = Slerp([ins["time"][i - 1], ins["time"][i]], Rotation.from_quat(ins["q"][i - 1:i + 1]))
        Rq = sl([t]).as_matrix()[0]
        p = ins["p"][i - 1] + (ins["p"][i] - ins["p"][i - 1]) * ((t - ins["time"][i - 1]) / (ins["time"][i] - ins["time"][i - 1]))
        assert np.allclose(R, Rq @ rig.R_b_c, atol=1e-12)
        assert np.allclose(pose[9:], p + Rq @ rig.t_b_c, atol=1e-12)
    # out of window: newest state, returns false (misc.cc:74-77)
    ok, pose = oracle.get_pose_from_ins_window(ins, rig.pose_b_c, ins["time"][-1] + 5.0)
    assert not ok
    Rl = Rotation.from_quat(ins["q"][-1]).as_matrix()
    assert np.allclose(pose[:9].reshape(3, 3).T, Rl @ rig.R_b_c, atol=1e-12)
    assert np.allclose(pose[9:], ins["p"][-1] + Rl @ rig.t_b_c, atol=1e-12)


# ----------------------------------------------------------------------------------------------
# projections (pointcloud.cc)
# ----------------------------------------------------------------------------------------------
def _pt(x, y, z, intensity=7.0):
    p = np.zeros(1, synth.XYZI)
    p["x"], p["y"], p["z"], p["w"], p["intensity"] = x, y, z, 1.0, intensity
    return p[0]


def test_point_projection_casts(oracle):
    rig = synth.rig_fusionportable()
    rng = np.random.default_rng(4)
    for _ in range(200):
        v = rng.normal(0, 30, 3).astype(np.float32)
        o = oracle.point_projection(rig.pose_c_l, _pt(*v))
        exp = (rig.R_c_l @ v.astype(np.float64) + rig.t_c_l).astype(np.float32)
        got = np.array([o["x"], o["y"], o["z"]])
        assert np.all(np.abs(got - exp) <= np.spacing(np.abs(exp)))  # identical up to the summation order's last ulp
        assert o["intensity"] == 7.0 and o["w"] == 1.0  # `projected = point` keeps the other fields (pointcloud.cc:66)


def test_relative_projection_to_camera_two_roundings(oracle):
    rig = synth.rig_r3live()
    ins = _ins(seed=5)
    _, pose0 = oracle.get_pose_from_ins_window(ins, rig.pose_b_c, ins["time"][50])
    _, pose1 = oracle.get_pose_from_ins_window(ins, rig.pose_b_c, ins["time"][20] + 1e-3)
    v = np.array([12.5, -3.25, 1.75], np.float32)
    o = oracle.relative_projection_to_camera(rig.pose_c_l, pose0, pose1, _pt(*v))
    R0, t0 = pose0[:9].reshape(3, 3).T, pose0[9:]
    R1, t1 = pose1[:9].reshape(3, 3).T, pose1[9:]
    c1 = (rig.R_c_l @ v.astype(np.float64) + rig.t_c_l).astype(np.float32)  # first f32 rounding (pointcloud.cc:38)
    exp = ((R0.T @ R1) @ c1.astype(np.float64) + R0.T @ (t1 - t0)).astype(np.float32)
    got = np.array([o["x"], o["y"], o["z"]])
    assert np.all(np.abs(got - exp) <= 2 * np.spacing(np.abs(exp)))
    # relativeProjection (no camera extrinsic)
    o2 = oracle.relative_projection(pose0, pose1, _pt(*v))
    exp2 = ((R0.T @ R1) @ v.astype(np.float64) + R0.T @ (t1 - t0)).astype(np.float32)
    assert np.all(np.abs(np.array([o2["x"], o2["y"], o2["z"]]) - exp2) <= 2 * np.spacing(np.abs(exp2)))
    assert oracle.lib().orc_point_square_distance is not None


# ----------------------------------------------------------------------------------------------
# ring buffer (le_vins.cc:149-163, 413-424)
# ----------------------------------------------------------------------------------------------
def test_ring_buffer_matches_deque_semantics(oracle):
    rng = np.random.default_rng(6)
    ring = oracle.Ring()
    dq = collections.deque()
    t = 500.0
    for scan in range(60):
        n = int(rng.integers(50, 400))
        pts = np.zeros(n, synth.XYZIT)
        pts["time"] = t + np.sort(rng.random(n)) * 0.3
        t += 0.3
        pts["x"] = scan
        ring.append(pts)
        dq.extend(pts["time"].tolist())
        while abs(dq[-1] - dq[0]) > 10.0 + 0.5 * 2:
            dq.popleft()
        assert len(ring) == len(dq)
    data = ring.data()
    assert np.array_equal(data["time"], np.array(dq))
    popped = ring.evict_before(t - 4.0)
    ref = sum(1 for v in dq if v < t - 4.0)
    assert popped == ref and len(ring) == len(dq) - ref


# ----------------------------------------------------------------------------------------------
# stage 1 against ground truth geometry
# ----------------------------------------------------------------------------------------------
def _surface_distance(p, boxes):
    d = []
    for lo, hi in boxes:
        inside = np.all((p >= lo) & (p <= hi), axis=1)
        dout = np.linalg.norm(np.maximum(np.maximum(lo - p, p - hi), 0), axis=1)
        din = np.min(np.minimum(p - lo, hi - p), axis=1)
        d.append(np.where(inside, din, dout))
    return np.min(np.stack(d), axis=0)


@pytest.mark.parametrize("cfg,pps", [(2, 4000), (3, None)])
def test_deskew_puts_points_back_on_the_scene(oracle, cfg, pps):
    """The generator ray-casts from the true sensor pose at every point's own time; a correct deskew therefore lands
    every kept point on a scene surface (1 cm range noise), while no deskew leaves decimetres of skew."""
    if cfg == 3:
        shp = synth.Shape(**{**synth.shape(3).__dict__, "stack_interval": 0.3})  # 3 scans, keeps the CPU suite short
        kf = synth.make_keyframe(3, shp=shp, n_scans_before=1)
    else:
        kf = synth.make_keyframe(cfg, pts_per_scan=pps)
    r = oracle.stack_deskew(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l,
                            kf.shape.stack_interval, details=True)
    assert r["ret"] == 1 and r["n"] > 0
    win = kf.points[r["first"]: r["first"] + r["n"]]
    # window membership (le_vins.cc:413-437)
    assert r["first"] == np.searchsorted(kf.points["time"], kf.stamp - kf.shape.stack_interval, side="left")
    assert r["first"] + r["n"] == np.searchsorted(kf.points["time"], kf.stamp, side="right")
    # range gate (le_vins.cc:448-451)
    norm = np.sqrt(win["x"] * win["x"] + (win["y"] * win["y"] + win["z"] * win["z"]), dtype=np.float32)
    kept = ~((norm > 200) | (norm < 1))
    assert np.array_equal(kept, r["kept"].astype(bool))
    c = r["cloud"]
    assert np.all(c["x"][~kept] == 0) and np.all(c["w"] == 1.0)
    # INS interval per point
    tt = win["time"][kept] + kf.rig.td_b_l
    assert np.array_equal(r["ins_index"][kept], np.searchsorted(kf.ins["time"], tt, side="right"))
    # geometry
    pc = np.stack([c["x"], c["y"], c["z"]], -1).astype(np.float64)[kept]
    Rwb, pwb = kf.shape.traj.pose(kf.stamp)
    Rwc, twc = Rwb @ kf.rig.R_b_c, pwb + Rwb @ kf.rig.t_b_c
    d = _surface_distance(pc @ Rwc.T + twc, kf.extra["boxes"])
    assert np.percentile(d, 99) < 0.05 and np.median(d) < 0.012
    raw = np.stack([win["x"], win["y"], win["z"]], -1).astype(np.float64)[kept]
    d0 = _surface_distance((raw @ kf.rig.R_c_l.T + kf.rig.t_c_l) @ Rwc.T + twc, kf.extra["boxes"])
    assert np.median(d0) > 5 * np.median(d)


def test_deskew_out_of_window_falls_back_to_newest_state(oracle):
    kf = synth.make_keyframe(2, pts_per_scan=2000)
    short = kf.ins[: int(len(kf.ins) * 0.85)]  # the INS window ends inside the stack window: misc.cc:74-77
    r = oracle.stack_deskew(kf.points, kf.stamp, short, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5, details=True)
    late = r["ins_index"] == 0
    assert late.any() and (r["ins_index"] > 0).any()
    # all fallback points share one rigid transform: the keyframe pose is the fallback pose too => camera frame = T_cl p
    win = kf.points[r["first"]: r["first"] + r["n"]][late]
    exp = (np.stack([win["x"], win["y"], win["z"]], -1).astype(np.float64) @ kf.rig.R_c_l.T + kf.rig.t_c_l)
    got = np.stack([r["cloud"]["x"], r["cloud"]["y"], r["cloud"]["z"]], -1)[late]
    assert np.allclose(got, exp, atol=2e-5)


def test_deskew_empty_cases(oracle):
    kf = synth.make_keyframe(2, pts_per_scan=1000)
    rig = kf.rig
    r = oracle.stack_deskew(kf.points, kf.stamp + 100.0, kf.ins, rig.pose_b_c, rig.pose_c_l, 0.0, 0.5)
    assert r["ret"] == -1 and r["n"] == 0  # everything popped: "No valid point cloud" (le_vins.cc:426-429)
    r = oracle.stack_deskew(kf.points, kf.points["time"][0] - 1.0, kf.ins, rig.pose_b_c, rig.pose_c_l, 0.0, 0.5)
    assert r["ret"] == 0 and r["n"] == 0 and r["first"] == 0  # nothing older than the stamp: empty cloud, returns false


# ----------------------------------------------------------------------------------------------
# stage 2 against a vectorised numpy restatement
# ----------------------------------------------------------------------------------------------
def _numpy_range_image(cloud, double_mode):
    x, y, z = cloud["x"], cloud["y"], cloud["z"]
    ang = 0.3 * np.pi / 180.0
    fov = 90.0 * np.pi / 180.0
    tan = np.tan(fov * 0.5)
    with np.errstate(divide="ignore", invalid="ignore"):
        ok = ~((z < 0) | (np.abs(x / z).astype(np.float64) > tan) | (np.abs(y / z).astype(np.float64) > tan))
        s = x * x + z * z
        if double_mode:
            ra = np.arctan2(y.astype(np.float64), np.sqrt(s.astype(np.float64))) + (np.pi * 0.5 - fov * 0.5)
            ca = np.arctan2(z.astype(np.float64), x.astype(np.float64)) - (np.pi * 0.5 - fov * 0.5)
        else:
            ra = np.arctan2(y, np.sqrt(s)).astype(np.float64) + (np.pi * 0.5 - fov * 0.5)
            ca = np.arctan2(z, x).astype(np.float64) - (np.pi * 0.5 - fov * 0.5)

    def rnd(v):  # C round(): half away from zero
        return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5))

    row = rnd(ra / ang).astype(np.int64)
    col = rnd(ca / ang).astype(np.int64)
    ok &= (row >= 0) & (col >= 0) & (row < 300) & (col < 300)
    cell = np.where(ok, col + row * 300, -1)
    dist = np.sqrt(x * x + (y * y + z * z))
    winner = np.full(90000, -1, np.int64)
    idx = np.nonzero(ok)[0]
    order = np.lexsort((idx, dist[idx], cell[idx]))  # per cell: nearest first, first index on ties
    srt = idx[order]
    first = np.ones(len(srt), bool)
    first[1:] = cell[srt][1:] != cell[srt][:-1]
    winner[cell[srt][first]] = srt[first]
    return cell, winner


@pytest.mark.parametrize("mode", [0, 1])
def test_range_image_vs_numpy(oracle, mode):
    kf = synth.make_keyframe(1, pts_per_scan=8000)
    r1 = oracle.stack_deskew(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5)
    cloud = r1["cloud"]
    ri = oracle.range_image(cloud, libm_mode=mode)
    cell, winner = _numpy_range_image(cloud, bool(mode))
    # numpy's own arctan2 loops may differ from glibc in the last ulp: allow a vanishing flip rate
    assert np.mean(ri["cell_of_point"] != cell) < 2e-4
    assert np.mean(ri["winner"] != winner) < 5e-4
    m = ri["m"]
    assert m == int(np.sum(ri["winner"] >= 0)) and m > 1000
    # compaction is row-major (tracking.cc:341-351) and the map is its inverse
    filled = np.nonzero(ri["winner"] >= 0)[0]
    assert np.array_equal(ri["cell_to_compact"][filled], np.arange(m))
    w = cloud[ri["winner"][filled]]
    v = np.stack([w["x"], w["y"], w["z"]], -1)
    n = np.sqrt(v[:, 0] * v[:, 0] + (v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]), dtype=np.float32)
    us = ri["unit_sphere"]
    assert np.array_equal(us["intensity"], n)
    assert np.array_equal(np.stack([us["x"], us["y"], us["z"]], -1), v / n[:, None])
    assert np.all(us["w"] == 1.0)


def test_range_image_edge_cases(oracle):
    empty = np.zeros(0, synth.XYZI)
    assert oracle.range_image(empty)["m"] == 0  # tracking.cc:356-358
    behind = np.zeros(100, synth.XYZI)
    behind["z"] = -5.0
    behind["x"] = 1.0
    assert oracle.range_image(behind)["m"] == 0  # all outside the FOV
    zeros = np.zeros(50, synth.XYZI)  # default points: 0/0 = NaN passes the FOV test, col = -150 fails
    zeros["w"] = 1.0
    r = oracle.range_image(zeros)
    assert r["m"] == 0 and np.all(r["cell_of_point"] == -1)
    # keep-nearest with first-index tie break (tracking.cc:332: strictly smaller replaces)
    c = np.zeros(4, synth.XYZI)
    c["z"] = [10.0, 5.0, 5.0, 7.0]
    r = oracle.range_image(c)
    assert r["m"] == 1 and r["winner"][r["winner"] >= 0][0] == 1
    assert r["unit_sphere"]["intensity"][0] == 5.0 and r["unit_sphere"]["z"][0] == 1.0


# ----------------------------------------------------------------------------------------------
# stage 3: k-NN, QR, plane fitting
# ----------------------------------------------------------------------------------------------
def _unit_cloud(m, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(0, 1, (m, 3))
    v[:, 2] = np.abs(v[:, 2]) + 1.5
    v /= np.linalg.norm(v, axis=1)[:, None]
    us = np.zeros(m, synth.XYZI)
    us["x"], us["y"], us["z"] = v.astype(np.float32).T
    us["w"] = 1.0
    us["intensity"] = rng.uniform(2, 60, m).astype(np.float32)
    return us


def test_knn_brute_vs_kdtree_vs_scipy_vs_flann(oracle):
    us = _unit_cloud(20000, 7)
    pts = np.stack([us["x"], us["y"], us["z"]], -1)
    tree = oracle.KdTree(us)
    ck = cKDTree(pts.astype(np.float64))
    rng = np.random.default_rng(8)
    q = pts[rng.integers(0, len(pts), 300)] + rng.normal(0, 2e-3, (300, 3)).astype(np.float32)
    flann = None
    try:
        import cv2

        flann = cv2.flann_Index(pts, {"algorithm": 4, "leaf_max_size": 15})
    except Exception:
        flann = None
    for k in range(len(q)):
        nb, ib, db = oracle.brute_knn(us, q[k])
        nk, ik, dk = tree.knn(q[k])
        assert nb == nk == 5
        assert np.array_equal(ib, ik) and np.array_equal(db, dk)
        assert np.all(np.diff(db) >= 0)
        _, isc = ck.query(q[k].astype(np.float64), 5)
        assert set(isc.tolist()) == set(ib.tolist())
        # L2_Simple accumulation order, f32
        d = q[k] - pts[ib]
        assert np.array_equal(db, (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        if flann is not None:
            fi, fd = flann.knnSearch(q[k:k + 1], 5, params={"checks": -1})
            assert set(fi[0].tolist()) == set(ib.tolist())
    n, idx, d2 = oracle.brute_knn(us[:3], q[0])
    assert n == 3 and idx[3] == -1  # fewer than 5 points exist


def test_colpiv_qr_vs_lstsq(oracle):
    rng = np.random.default_rng(9)
    for _ in range(300):
        A = rng.normal(0, 1, (5, 3)) * rng.choice([1.0, 30.0]) + rng.normal(0, 5, 3)
        b = -np.ones(5)
        x = oracle.colpiv_qr_solve_5x3(A, b)
        ref = np.linalg.lstsq(A, b, rcond=None)[0]
        assert np.allclose(x, ref, rtol=1e-8, atol=1e-10)
    # nearly planar depth-association geometry (cond ~ 1e3)
    n = np.array([0.1, -0.2, 0.97])
    n /= np.linalg.norm(n)
    P = rng.normal(0, 0.05, (5, 3)) + np.array([0.3, 0.2, 12.0])
    P -= np.outer(P @ n - 11.0, n)  # on the plane n.p = 11
    x = oracle.colpiv_qr_solve_5x3(P, -np.ones(5))
    assert np.allclose(x / np.linalg.norm(x), -n, atol=1e-9) and np.isclose(1 / np.linalg.norm(x), 11.0, rtol=1e-9)
    # rank deficient: all rows equal -> minimum-pivot solution, finite
    x = oracle.colpiv_qr_solve_5x3(np.tile([1.0, 2.0, 3.0], (5, 1)), -np.ones(5))
    assert np.all(np.isfinite(x)) and np.isclose(np.dot(x, [1, 2, 3]), -1.0)


def _plane_image(normal, d, n_side=41, spacing=0.3, noise=0.0, seed=0):
    """A dense fronto-ish plane n.p = d seen through a regular 0.3-degree grid of rays."""
    rng = np.random.default_rng(seed)
    a = np.deg2rad((np.arange(n_side) - n_side // 2) * spacing)
    el, az = np.meshgrid(a, a, indexing="ij")
    rays = np.stack([np.cos(el) * np.sin(az), np.sin(el), np.cos(el) * np.cos(az)], -1).reshape(-1, 3)
    t = d / (rays @ normal) + rng.normal(0, noise, len(rays))
    cloud = np.zeros(len(rays), synth.XYZI)
    p = (rays * t[:, None]).astype(np.float32)
    cloud["x"], cloud["y"], cloud["z"], cloud["w"] = p[:, 0], p[:, 1], p[:, 2], 1.0
    return cloud


def test_plane_fitting_recovers_analytic_depth(oracle):
    normal = np.array([0.15, -0.1, 1.0])
    normal /= np.linalg.norm(normal)
    cloud = _plane_image(normal, 8.0, noise=0.002)
    ri = oracle.range_image(cloud)
    rng = np.random.default_rng(10)
    q = np.stack([rng.uniform(-0.05, 0.05, 50), rng.uniform(-0.05, 0.05, 50), np.ones(50)], -1)
    q = (q / np.linalg.norm(q, axis=1)[:, None]).astype(np.float32)
    a = oracle.associate(ri["unit_sphere"], q)
    assert np.all(a["type"] == 1) and np.all(a["flags"] == 0)
    t_true = 8.0 / (q.astype(np.float64) @ normal)
    assert np.allclose(a["xyz_depth"][:, :3], q * t_true[:, None], atol=0.02)
    assert np.array_equal(a["xyz_depth"][:, 3], a["xyz_depth"][:, 2])  # depth = z (tracking.cc:629)
    assert all(oracle.is_good_depth(d) for d in a["xyz_depth"][:, 3])
    b = oracle.associate(ri["unit_sphere"], q, use_kdtree=True)
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_plane_fitting_rejections(oracle):
    normal = np.array([0.0, 0.0, 1.0])
    q0 = np.array([[0.0, 0.0, 1.0]], np.float32)
    # 5th neighbour too far: a sparse image (every 3rd ray)  (tracking.cc:573-575)
    sparse = _plane_image(normal, 8.0, n_side=15, spacing=1.2)
    a = oracle.associate(oracle.range_image(sparse)["unit_sphere"], q0)
    assert a["type"][0] == -1 and a["flags"][0] == 2
    assert np.array_equal(a["xyz_depth"][0], [0, 0, 1, 0])  # the query is left untouched, depth 0
    # fewer than 5 points (reference: out-of-bounds read; build: NONE)
    a = oracle.associate(oracle.range_image(sparse[:3])["unit_sphere"], q0)
    assert a["type"][0] == -1 and a["flags"][0] == 1
    # range spread > 0.5 m between the nearest and the 5th neighbour (tracking.cc:607-611): hand-made neighbourhood
    offs = np.deg2rad([[0.05, 0.0], [0.0, 0.1], [-0.15, 0.0], [0.0, -0.2], [0.25, 0.0]])
    us = np.zeros(5, synth.XYZI)
    d = np.stack([np.sin(offs[:, 0]), np.sin(offs[:, 1]), np.ones(5)], -1)
    d /= np.linalg.norm(d, axis=1)[:, None]
    us["x"], us["y"], us["z"], us["w"] = d[:, 0].astype(np.float32), d[:, 1].astype(np.float32), d[:, 2].astype(np.float32), 1.0
    us["intensity"] = np.array([10.0, 9.9, 9.8, 9.7, 9.3], np.float32)
    a = oracle.associate(us, q0)
    assert a["type"][0] == -1 and a["flags"][0] & 8 and np.array_equal(a["knn_idx"][0], [0, 1, 2, 3, 4])
    us["intensity"] = np.array([10.0, 9.9, 9.8, 9.7, 9.6], np.float32)
    assert not (oracle.associate(us, q0)["flags"][0] & 8)
    # plane residual > 0.1 m: rough surface with small range spread
    rough = _plane_image(normal, 40.0, noise=0.15, seed=3)
    r = oracle.associate(oracle.range_image(rough)["unit_sphere"], np.repeat(q0, 1, 0))
    assert r["flags"][0] & (4 | 8) or r["type"][0] == 1
    # t outside [1, 200] (tracking.cc:621-623): a valid plane closer than 1 m
    near = _plane_image(normal, 0.6)
    a = oracle.associate(oracle.range_image(near)["unit_sphere"], q0)
    assert a["type"][0] == -1 and a["flags"][0] == 16
    # t >= 1 but z <= 1: passes plane fitting, rejected by isGoodDepth (SURVEY 8a note 7)
    assert not oracle.is_good_depth(1.0) and oracle.is_good_depth(1.0001) and not oracle.is_good_depth(200.0)


def test_plane_estimation_uncalled_surface(oracle):
    normal = np.array([0.2, 0.1, 1.0])
    normal /= np.linalg.norm(normal)
    cloud = _plane_image(normal, 5.0, n_side=3)
    ok, n, ninv, dist = oracle.plane_estimation(cloud[:5], 0.05)
    assert ok and len(dist) == 5 and np.allclose(np.abs(n), np.abs(normal), atol=1e-5) and np.isclose(ninv, 5.0, rtol=1e-5)
    bad = cloud[[0, 2, 4, 6, 8]].copy()  # four corners and the centre of the 3 x 3 patch
    assert oracle.plane_estimation(bad, 0.05)[0]
    bad["z"][2] += 1.0  # the centre pushed off the plane cannot be absorbed by tilting it
    ok, *_ = oracle.plane_estimation(bad, 0.05)
    assert not ok
    assert not oracle.plane_estimation(cloud[:4], 0.05)[0]


def test_camera_functions(oracle):
    intr = np.array([863.4241, 863.4171, 640.6808, 518.3392, 0.0])
    pc = oracle.pixel2cam(intr, 100.5, 700.25)
    assert np.allclose(pc, [(np.float32(100.5) - intr[2]) / intr[0], (np.float32(700.25) - intr[3]) / intr[1], 1.0])
    puc = oracle.pixel2unitcam(intr, 100.5, 700.25)
    assert np.isclose(np.linalg.norm(puc), 1.0) and np.allclose(puc, pc / np.linalg.norm(pc))
    rays = synth.pixels_to_unit_rays(intr, np.array([[100.5, 700.25]], np.float32))
    assert np.array_equal(rays[0], puc.astype(np.float32))
    rig = synth.rig_r3live()
    assert np.allclose(oracle.cam2world([1.0, 2.0, 3.0], rig.pose_b_c), rig.R_b_c @ [1, 2, 3] + rig.t_b_c)


def test_whole_keyframe_driver_consistency(oracle):
    kf = synth.make_keyframe(2, pts_per_scan=3000, n_features=80)
    whole = oracle.keyframe(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries, 0.5)
    r1 = oracle.stack_deskew(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5)
    r2 = oracle.range_image(r1["cloud"])
    r3 = oracle.associate(r2["unit_sphere"], kf.queries)
    assert whole["m"] == r2["m"] and whole["n_projected"] == r1["n"]
    assert np.array_equal(whole["type"], r3["type"]) and np.array_equal(whole["xyz_depth"], r3["xyz_depth"])
    assert (whole["type"] == 1).sum() > 10


def test_undistort_points_equals_opencv(oracle):
    """orc_undistort_points restates cv::undistortPoints(pts, pts, K, D, Mat(), K) (camera.cc:72-74): bit for bit against the
    OpenCV in this image, with radial + tangential coefficients, with k3 and with a skewed K."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    cases = [((863.4241, 863.4171, 640.6808, 518.3392, 0.0), (-0.1080, 0.1050, -1.2872e-4, 5.7923e-5, -0.0222)),
             ((593.5669, 593.8849, 513.3945, 401.9319, 0.0), (-0.33, 0.11, 0.0005, -0.0003, 0.0)),
             ((458.654, 457.296, 367.215, 248.375, 0.7), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)),
             ((458.654, 457.296, 367.215, 248.375, 0.0), (0.0, 0.0, 0.0, 0.0, 0.0))]
    for intr, dist in cases:
        px = np.stack([rng.uniform(-50, 1400, 30000), rng.uniform(-50, 1100, 30000)], -1).astype(np.float32)
        K = np.array([[intr[0], intr[4], intr[2]], [0, intr[1], intr[3]], [0, 0, 1]])
        ref = cv2.undistortPoints(px.reshape(-1, 1, 2), K, np.array(dist).reshape(5, 1), None, K).reshape(-1, 2)
        got = oracle.undistort_points(intr, dist, px)
        assert got.tobytes() == ref.tobytes()


def test_depth_factor_residual_jacobian_and_huber_cost(oracle):
    """DepthFactor (depth_factor.h:38-42): r = (depth - 1 / invdepth) / std, AutoDiff Jacobian dr / d invdepth = 1 / (invdepth^2 std);
    the loss at its call site (optimizer.cc:659-665) is Ceres' HuberLoss(a): rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 beyond,
    cost = rho(r^2) / 2.  Checked against the closed forms, a central difference and the continuity of rho at s = a^2."""
    rng = np.random.default_rng(21)
    for _ in range(400):
        depth = rng.uniform(1.0, 200.0)
        invd = 1.0 / rng.uniform(0.5, 250.0)
        sigma = rng.choice([0.1, 0.05, 1.0])
        a = rng.choice([0.0, 1.0, 2.5])
        r, j, c = oracle.depth_factor(depth, invd, sigma, a)
        assert r == (depth - 1.0 / invd) / sigma
        assert j == (1.0 / (invd * invd)) / sigma
        h = invd * 1e-6
        rp, _, _ = oracle.depth_factor(depth, invd + h, sigma, a)
        rm, _, _ = oracle.depth_factor(depth, invd - h, sigma, a)
        assert abs((rp - rm) / (2 * h) - j) <= 1e-6 * abs(j)
        s = r * r
        rho = s if (a <= 0 or s <= a * a) else 2.0 * a * np.sqrt(s) - a * a
        assert c == 0.5 * rho
        assert c <= 0.5 * s + 1e-12  # Huber never exceeds the squared loss
    # continuity at the corner s = a^2 and linear growth beyond it
    a, sigma, invd = 1.0, 0.1, 0.1
    for eps in (-1e-9, 0.0, 1e-9):
        r, _, c = oracle.depth_factor(10.0 + (a + eps) * sigma, invd, sigma, a)
        assert abs(c - 0.5 * a * a) < 1e-7
    _, _, c2 = oracle.depth_factor(10.0 + 3.0 * sigma, invd, sigma, a)
    _, _, c4 = oracle.depth_factor(10.0 + 5.0 * sigma, invd, sigma, a)
    assert abs((c4 - c2) - a * 2.0) < 1e-6  # d cost / d|r| = a beyond the corner


def test_chi2_outlier_semantics(oracle):
    """optimizer.cc:752-783: cost = r^T r of the 2-vector residual, outlier iff cost > 5.991 (strictly)."""
    rng = np.random.default_rng(22)
    for _ in range(300):
        r = rng.normal(size=2) * 2.0
        out, cost = oracle.chi2_outlier(r, 5.991)
        assert cost == r[0] * r[0] + r[1] * r[1]
        assert out == (cost > 5.991)
    x = np.sqrt(5.991 / 2.0)
    out, cost = oracle.chi2_outlier(np.array([x, x]), cost_probe := float(x * x + x * x))
    assert not out and cost == cost_probe  # equal to the threshold is not an outlier
    assert oracle.chi2_outlier(np.array([x, np.nextafter(x, 10.0)]), cost_probe)[0]
    assert oracle.chi2_outlier(np.array([0.0, 0.0]))[0] is False


def _point_to_plane_std_numpy(pts, target, point_std):
    """pointcloud.cc:111-190 restated with numpy (eigh instead of Eigen::EigenSolver; the formula does not depend on the
    signs or the order of the eigenvectors)"""
    P = np.stack([pts["x"], pts["y"], pts["z"]], -1).astype(np.float64)
    t = np.array([target["x"], target["y"], target["z"]], np.float64)
    n_pts = len(P)
    inv = 1.0 / n_pts
    var_p = np.eye(3) * point_std * point_std
    center = P.sum(0) * inv
    cov = (P.T @ P) * inv - np.outer(center, center)
    w, V = np.linalg.eigh(cov)
    i_min = int(np.argmin(w))
    nv = V[:, i_min]
    plane_cov = np.zeros((6, 6))
    for pv in P:
        F = np.zeros((3, 3))
        for m in range(3):
            if m != i_min:
                F[m] = inv * (pv - center) / (w[i_min] - w[m]) @ (np.outer(V[:, m], nv) + np.outer(nv, V[:, m]))
        J = np.zeros((6, 3))
        J[:3] = V @ F
        J[3:] = np.eye(3) * inv
        plane_cov += J @ var_p @ J.T
    j = np.concatenate([t - center, -nv])
    return float(np.sqrt(j @ plane_cov @ j + nv @ var_p @ nv))


def test_point_to_plane_std_vs_numpy(oracle):
    rng = np.random.default_rng(33)
    XYZI = oracle.XYZI
    worst = 0.0
    for trial in range(300):
        normal = rng.normal(size=3)
        normal /= np.linalg.norm(normal)
        a = np.cross(normal, [1.0, 0.3, -0.2])
        a /= np.linalg.norm(a)
        b = np.cross(normal, a)
        n_pts = 5 if trial % 3 else int(rng.integers(4, 12))
        c = rng.uniform(-20, 20, 3)
        uv = rng.uniform(-1.0, 1.0, (n_pts, 2)) * [1.0, 0.4]  # anisotropic patch: three well separated eigenvalues
        P = c + uv[:, :1] * a + uv[:, 1:] * b + rng.normal(size=(n_pts, 1)) * 0.01 * normal
        pts = np.zeros(n_pts, XYZI)
        pts["x"], pts["y"], pts["z"] = P[:, 0], P[:, 1], P[:, 2]
        tgt = np.zeros(1, XYZI)
        tp = c + rng.uniform(-1, 1) * a + rng.uniform(-0.4, 0.4) * b + rng.normal() * 0.05 * normal
        tgt["x"], tgt["y"], tgt["z"] = tp
        got = oracle.point_to_plane_std(pts, tgt[0], 0.03)
        want = _point_to_plane_std_numpy(pts, tgt[0], 0.03)
        assert np.isfinite(got) and got >= 0.03 * (1 - 1e-9)  # never below the target point's own sigma along the normal
        worst = max(worst, abs(got - want) / want)
    assert worst < 1e-7, worst
