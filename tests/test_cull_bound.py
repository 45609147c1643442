"""The certified frustum cull of k_deskew_bin, restated in numpy and checked on the CPU against the oracle:
whenever the f32 test `certainly_outside` fires for a point, the reference's FOV gate (tracking.cc:311-314) rejects that point
after the exact f64 deskew.  This validates the bound of `fill_cull_record` (csrc/kernels.cuh) independently of the GPU tests,
which check the kernel's results (tests/test_gpu_edges.py::test_frustum_cull_is_exact_on_the_fov_border, the fuzz tests)."""
import numpy as np
import pytest

from levins import synth


def _rod(k, phi):
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(phi) * K + (1 - np.cos(phi)) * (K @ K)


def cull_record(oracle, s0, s1, R0, t0, R_bc, t_bc, R_cl, t_cl):
    """numpy twin of fill_cull_record for the interval [s0, s1] and the keyframe pose (R0, t0)"""
    q0 = s0["q"] / np.linalg.norm(s0["q"])
    Rq0 = oracle.quat_to_matrix(q0)
    qi = np.array([-s1["q"][0], -s1["q"][1], -s1["q"][2], s1["q"][3]]) / np.dot(s1["q"], s1["q"])
    a, b = qi, s0["q"]  # dq = q1^-1 q0 (x, y, z, w)
    dq = np.array([a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
                   a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
                   a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0],
                   a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]])
    rv = np.array(oracle.quaternion2vector(dq))
    theta = np.linalg.norm(rv)
    k = rv / theta if theta > 0 else np.array([1.0, 0, 0])
    A = R0.T @ Rq0
    c0 = R0.T @ (s0["p"] - t0)
    c1 = R0.T @ (s1["p"] - s0["p"])
    inv_dts = 1.0 / (s1["state_time"] - s0["state_time"])
    s_lo, s_hi = (s0["time"] - s0["state_time"]) * inv_dts, (s1["time"] - s0["state_time"]) * inv_dts
    s_mid, h = 0.5 * (s_lo + s_hi), 0.5 * abs(s_hi - s_lo) * 1.001 + 1e-9
    AR = A @ _rod(k, -s_mid * theta)
    M = AR @ R_bc @ R_cl
    m = AR @ (R_bc @ t_cl + t_bc) + c0 + s_mid * c1
    fro = np.linalg.norm
    nk = fro(k)
    rot = fro(A) * h * abs(theta) * (nk + nk * nk)
    ca = (rot * fro(R_bc) * fro(R_cl) + 1e-5 * fro(M)) * 1.01
    cb = (rot * (fro(R_bc) * fro(t_cl) + fro(t_bc)) + h * fro(c1) + 1e-5 * fro(m) + 1e-6) * 1.01
    return M.astype(np.float32), m.astype(np.float32), np.float32(ca) * np.float32(1.000001), np.float32(cb) * np.float32(1.000001)


def certainly_outside(M, m, ca, cb, p):
    f = np.float32
    pt = (M @ p.T).T.astype(f) + m  # (the kernel's FMA chain differs from this by roundings the bound covers)
    n = np.sqrt((p.astype(f) ** 2).sum(axis=1)).astype(f)
    d = ca * n + cb
    return (pt[:, 2] < -d) | (np.maximum(np.abs(pt[:, 0]), np.abs(pt[:, 1])) - np.abs(pt[:, 2]) > f(2.001) * d)


@pytest.mark.parametrize("seed", range(6))
def test_cull_never_drops_a_point_the_reference_keeps(oracle, seed):
    rng = np.random.default_rng(seed)
    rig = synth.rig_fusionportable() if seed % 2 else synth.rig_r3live()
    rate = [200.0, 100.0, 50.0][seed % 3]
    t = 345600.0 + np.arange(60) / rate
    ins = np.zeros(len(t), synth.INS)
    ins["time"] = t
    ins["state_time"] = t + (1e-4 if seed % 2 else 0.0)
    w = rng.normal(size=3) * [0.05, 0.3, 0.25][seed % 3]  # rad/s: theta per interval stays below 2^-6 ("fast" entries)
    v = rng.normal(size=3) * [0.5, 5.0, 15.0][seed % 3]
    ang = np.linalg.norm(w) * (t - t[0])
    ins["q"][:, :3] = np.sin(0.5 * ang)[:, None] * (w / np.linalg.norm(w))
    ins["q"][:, 3] = np.cos(0.5 * ang)
    ins["p"] = v[None, :] * (t - t[0])[:, None]
    stamp = float(t[40] + 0.3 / rate)
    ok, pose0 = oracle.get_pose_from_ins_window(ins, rig.pose_b_c, stamp)
    R0, t0 = pose0[:9].reshape(3, 3).T, pose0[9:]
    n_culled = n_total = 0
    for j in rng.choice(np.arange(1, 40), 8, replace=False):
        M, m, ca, cb = cull_record(oracle, ins[j - 1], ins[j], R0, t0, rig.R_b_c, rig.t_b_c, rig.R_c_l, rig.t_c_l)
        assert 0 <= ca < 0.05
        n = 4000
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = (d * rng.uniform(1.0, 150.0, (n, 1))).astype(np.float32)
        pts = np.zeros(n, synth.XYZIT)
        pts["x"], pts["y"], pts["z"] = p[:, 0], p[:, 1], p[:, 2]
        # INS times spread over the whole interval (both ends included)
        pts["time"] = np.sort(np.concatenate([[ins["time"][j - 1]], rng.uniform(ins["time"][j - 1], ins["time"][j], n - 2),
                                              [np.nextafter(ins["time"][j], 0)]])) - rig.td_b_l
        # the reference's path: exact deskew of every point (window logic bypassed: stack interval covers everything)
        ref = oracle.stack_deskew(pts, float(pts["time"][-1]) + 1e-6, ins, rig.pose_b_c, rig.pose_c_l, rig.td_b_l, 1e6, details=True)
        assert ref["n"] == n
        # (the oracle deskews into the frame at ITS stamp; use the same stamp for the keyframe pose)
        ok2, pose0b = oracle.get_pose_from_ins_window(ins, rig.pose_b_c, float(pts["time"][-1]) + 1e-6)
        R0b, t0b = pose0b[:9].reshape(3, 3).T, pose0b[9:]
        M, m, ca, cb = cull_record(oracle, ins[j - 1], ins[j], R0b, t0b, rig.R_b_c, rig.t_b_c, rig.R_c_l, rig.t_c_l)
        cells = oracle.range_image(ref["cloud"])  # cell ids of the exact path
        kept_by_reference = np.zeros(n, bool)
        win = cells["winner"]
        # a point is kept by the FOV gate iff it has a cell; recompute the gate directly from the exact coordinates
        c = ref["cloud"]
        x, y, z = c["x"], c["y"], c["z"]
        with np.errstate(divide="ignore", invalid="ignore"):
            tan = np.tan(np.pi / 4)
            rejected = (z < 0) | (np.abs((x / z).astype(np.float32)).astype(np.float64) > tan) | (np.abs((y / z).astype(np.float32)).astype(np.float64) > tan)
        kept_by_reference = ~rejected & ref["kept"].astype(bool)
        out = certainly_outside(M, m, ca, cb, p)
        assert not np.any(out & kept_by_reference), "the cull fired for a point the reference keeps"
        # and the bound itself: |p~ - p_c0| <= ca |p| + cb on every range-gated point
        gated = ref["kept"].astype(bool)
        pt = (M.astype(np.float64) @ p.T.astype(np.float64)).T + m
        err = np.linalg.norm(pt - np.stack([x, y, z], -1).astype(np.float64), axis=1)
        bound = float(ca) * np.linalg.norm(p.astype(np.float64), axis=1) + float(cb)
        assert np.all(err[gated] <= bound[gated]), float((err[gated] / bound[gated]).max())
        n_culled += int(out.sum())
        n_total += n
        assert win is not None
    assert n_culled > 0.55 * n_total  # isotropic directions: ~5/6 lie outside the 90-degree cone, most of them are culled


def interval_entry(oracle, s0, s1, R0, t0):
    """numpy (f64) twin of the interval entry k_prepare builds (kernels.cuh): A, axis k, angle theta, c0, c1, ts0, 1 / dts"""
    q0 = s0["q"] / np.linalg.norm(s0["q"])
    a = np.array([-s1["q"][0], -s1["q"][1], -s1["q"][2], s1["q"][3]]) / np.dot(s1["q"], s1["q"])
    b = s0["q"]
    dq = np.array([a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
                   a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
                   a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0],
                   a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]])
    rv = np.array(oracle.quaternion2vector(dq))
    theta = np.linalg.norm(rv)
    k = rv / theta if theta > 0 else np.array([1.0, 0, 0])
    return (R0.T @ oracle.quat_to_matrix(q0), k, theta, R0.T @ (s0["p"] - t0), R0.T @ (s1["p"] - s0["p"]),
            float(s0["state_time"]), 1.0 / float(s1["state_time"] - s0["state_time"]))


@pytest.mark.parametrize("seed", range(4))
def test_interval_entry_algebra_reproduces_the_exact_deskew(oracle, seed):
    """DESIGN.md 3, "Stage 1 algebra": per point the kernels evaluate  f32(A Rod(k, -s theta) (R_bc f32(R_cl p + t_cl) + t_bc)
    + c0 + s c1)  with per-interval constants instead of the reference's interpolate-compose-project chain (misc.cc:80-103,
    pointcloud.cc:36-48,65-71).  Evaluated here in numpy f64 with both f32 roundings, it gives the oracle's deskewed
    coordinates bit for bit on > 99.9 % of the points and within one f32 ulp on the rest (the two orders of operations differ
    by ~1e-16 relative before the last rounding, which only shows on a rounding boundary)."""
    rng = np.random.default_rng(100 + seed)
    rig = synth.rig_fusionportable() if seed % 2 else synth.rig_r3live()
    rate = [200.0, 100.0][seed % 2]
    t = 345600.0 + np.arange(60) / rate
    ins = np.zeros(len(t), synth.INS)
    ins["time"] = t
    ins["state_time"] = t + (1e-4 if seed % 2 else 0.0)
    w = rng.normal(size=3) * [0.3, 2.0, 0.05, 6.0][seed]  # rad/s; the last one leaves the small-angle tier (theta > 2^-6)
    v = rng.normal(size=3) * [5.0, 15.0, 0.5, 3.0][seed]
    ang = np.linalg.norm(w) * (t - t[0])
    ins["q"][:, :3] = np.sin(0.5 * ang)[:, None] * (w / np.linalg.norm(w))
    ins["q"][:, 3] = np.cos(0.5 * ang)
    ins["p"] = v[None, :] * (t - t[0])[:, None] + 0.5 * rng.normal(size=3)[None, :] * ((t - t[0]) ** 2)[:, None]
    f32, f64 = np.float32, np.float64
    R_bc, t_bc, R_cl, t_cl = rig.R_b_c, rig.t_b_c, rig.R_c_l, rig.t_c_l
    n_same = n_all = 0
    for j in rng.choice(np.arange(1, 40), 6, replace=False):
        n = 20000
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = (d * rng.uniform(1.0, 150.0, (n, 1))).astype(f32)
        pts = np.zeros(n, synth.XYZIT)
        pts["x"], pts["y"], pts["z"] = p[:, 0], p[:, 1], p[:, 2]
        pts["time"] = np.sort(np.concatenate([[ins["time"][j - 1]], rng.uniform(ins["time"][j - 1], ins["time"][j], n - 2),
                                              [np.nextafter(ins["time"][j], 0)]])) - rig.td_b_l
        stamp = float(pts["time"][-1]) + 1e-6
        ref = oracle.stack_deskew(pts, stamp, ins, rig.pose_b_c, rig.pose_c_l, rig.td_b_l, 1e6, details=True)
        assert ref["n"] == n
        ok, pose0 = oracle.get_pose_from_ins_window(ins, rig.pose_b_c, stamp)
        R0, t0 = pose0[:9].reshape(3, 3).T, pose0[9:]
        A, k, theta, c0, c1, ts0, inv_dts = interval_entry(oracle, ins[j - 1], ins[j], R0, t0)
        tt = pts["time"].astype(f64) + rig.td_b_l
        s = (tt - ts0) * inv_dts
        p1 = ((R_cl @ p.astype(f64).T).T + t_cl).astype(f32)  # first rounding (pointcloud.cc:38 -> 68)
        u = (R_bc @ p1.astype(f64).T).T + t_bc
        phi = -s * theta
        ku = np.cross(k[None, :], u)
        kku = np.cross(k[None, :], ku)
        wv = u + np.sin(phi)[:, None] * ku + (2.0 * np.sin(0.5 * phi) ** 2)[:, None] * kku
        out = ((A @ wv.T).T + c0[None, :] + s[:, None] * c1[None, :]).astype(f32)  # second rounding (pointcloud.cc:45 -> 68)
        gated = ref["kept"].astype(bool)
        want = np.stack([ref["cloud"]["x"], ref["cloud"]["y"], ref["cloud"]["z"]], -1)[gated]
        got = out[gated]
        same = got.view(np.uint32) == want.view(np.uint32)
        ulp = np.spacing(np.maximum(np.abs(want), f32(1e-30)))
        assert np.all(np.abs(got.astype(f64) - want.astype(f64)) <= ulp.astype(f64)), "more than one f32 ulp from the oracle"
        n_same += int(same.all(axis=1).sum())
        n_all += int(gated.sum())
    assert n_all > 100_000 and n_same >= 0.999 * n_all, (n_same, n_all)
