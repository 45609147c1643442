"""driver records -> PointXYZIT (ROS/lidar/lidar_converter.cc:35-256).  CPU: the oracle's restatement
against an independent numpy one for the four drivers.  GPU: levins_append_raw against the oracle, bit for bit, and a
keyframe fed with driver records against one fed with converted points."""
import numpy as np
import pytest

from levins import rawfmt, synth

WEEK = 2334


def _np_unix2gps(u):
    seconds = u + 18 - 315964800
    week = np.floor(seconds / 604800).astype(np.int64)
    return seconds - (week * 604800).astype(np.float64)


def _numpy_convert(kind, rec, timebase_ns, stamp, to_gps, scan_line, nearest, farthest):
    x, y, z = rec["x"].astype(np.float32), rec["y"].astype(np.float32), rec["z"].astype(np.float32)
    if kind == "livox":
        ok = (rec["line"] < scan_line) & (((rec["tag"] & 0x30) == 0) | ((rec["tag"] & 0x30) == 0x10))
        t = (np.uint64(timebase_ns) + rec["offset_time"].astype(np.uint64)).astype(np.float64) * 1.0e-9
        t = _np_unix2gps(t) if to_gps else t
        inten = rec["reflectivity"].astype(np.float32)
    else:
        ok = np.ones(len(rec), bool)
        inten = rec["intensity"]
        if kind == "hesai":
            t = _np_unix2gps(rec["timestamp"]) if to_gps else rec["timestamp"].copy()
        else:
            s = float(_np_unix2gps(np.float64(stamp))) if to_gps else stamp
            t = s + rec["time"].astype(np.float64) if kind == "velodyne" else s + rec["t"].astype(np.float64) * 1.0e-9
    sq = (x * x + (y * y + z * z)).astype(np.float64)
    keep = ok & (sq > nearest * nearest) & (sq < farthest * farthest)
    out = np.zeros(int(keep.sum()), synth.XYZIT)
    out["x"], out["y"], out["z"], out["intensity"], out["time"] = x[keep], y[keep], z[keep], inten[keep], t[keep]
    return out, (t[ok].min() if ok.any() else np.finfo(np.float64).max), max(0.0, t[ok].max()) if ok.any() else 0.0


def _cases(seed=0, n=30_000):
    rng = np.random.default_rng(seed)
    kf = synth.make_keyframe(2, pts_per_scan=n // 5)
    pts = kf.points[:n].copy()
    pts["x"][rng.random(len(pts)) < 0.02] = 0.0  # invalid returns: (0, 0, 0) and out-of-range points
    far = rng.random(len(pts)) < 0.01
    pts["z"][far] = 500.0
    unix = pts["time"] + WEEK * 604800 + rawfmt.GPS_UNIX_OFFSET
    stamp = float(unix.min())
    liv, tb = rawfmt.livox_records(pts, WEEK, rng=rng, bad_fraction=0.05)
    vel = np.zeros(len(pts), rawfmt.VELODYNE)
    ous = np.zeros(len(pts), rawfmt.OUSTER)
    hes = np.zeros(len(pts), rawfmt.HESAI)
    for r in (vel, ous, hes):
        r["x"], r["y"], r["z"], r["intensity"] = pts["x"], pts["y"], pts["z"], pts["intensity"]
        r["ring"] = rng.integers(0, 16, len(pts))
    vel["time"] = (unix - stamp).astype(np.float32)
    ous["t"] = np.round((unix - stamp) * 1e9).astype(np.uint32)
    hes["timestamp"] = unix
    return [("livox", rawfmt.fmt_livox(), liv, tb, 0.0, 6), ("livox19", rawfmt.fmt_livox_wire(), liv, tb, 0.0, 6),
            ("velodyne", rawfmt.fmt_velodyne(), vel, 0, stamp, 16), ("ouster", rawfmt.fmt_ouster(), ous, 0, stamp, 128),
            ("hesai", rawfmt.fmt_hesai(), hes, 0, stamp, 32)]


def _bytes(name, rec):
    return rawfmt.as_bytes(rec, 19 if name == "livox19" else None)


@pytest.mark.parametrize("to_gps", [True, False])
def test_oracle_converter_against_numpy(oracle, to_gps):
    for name, fmt, rec, tb, stamp, lines in _cases():
        got, st, en = oracle.convert_raw(fmt, _bytes(name, rec), len(rec), tb, stamp, to_gps, lines, 1.0, 200.0)
        want, wst, wen = _numpy_convert(name.rstrip("19"), rec, tb, stamp, to_gps, lines, 1.0, 200.0)
        assert len(got) == len(want) and 0 < len(got) < len(rec), name
        assert got.tobytes() == want.tobytes(), name
        assert st == wst and en == wen, name
    assert oracle.unix2gps_sow(1.7e9) == float(_np_unix2gps(np.float64(1.7e9)))


@pytest.mark.gpu
def test_append_raw_against_oracle(ctx, oracle):
    from levins import abi

    for name, fmt, rec, tb, stamp, lines in _cases(seed=3):
        raw = _bytes(name, rec)
        want, wst, wen = oracle.convert_raw(fmt, raw, len(rec), tb, stamp, True, lines, 1.0, 200.0)
        ctx.buffer_clear()
        m, st, en = ctx.append_raw(fmt, raw, len(rec), timebase_ns=tb, stamp=stamp, to_gps_time=True, scan_line=lines)
        assert m == len(want) and st == wst and en == wen, name
        size, t_front, t_back = ctx.buffer_info()
        assert size == m and t_front == want["time"][0] and t_back == want["time"][-1]
    # empty and fully rejected scans
    name, fmt, rec, tb, stamp, lines = _cases(seed=3)[0]
    ctx.buffer_clear()
    assert ctx.append_raw(fmt, _bytes(name, rec[:0]), 0, timebase_ns=tb, scan_line=lines)[0] == 0
    m, st, en = ctx.append_raw(fmt, _bytes(name, rec), len(rec), timebase_ns=tb, scan_line=0)  # no line passes
    assert m == 0 and st == np.finfo(np.float64).max and en == 0.0
    with pytest.raises(abi.LevinsError):
        bad = fmt.copy()
        bad["off_time"] = 40
        ctx.append_raw(bad, _bytes(name, rec), len(rec))


@pytest.mark.gpu
def test_keyframe_from_driver_records_equals_keyframe_from_points(ctx, oracle):
    """Scans arrive as Livox records (wrapping a small ring on the way); the keyframe must equal the one computed from the
    oracle-converted points, which in turn equals the oracle's keyframe."""
    from levins import abi

    kf = synth.make_keyframe(2, pts_per_scan=6000, n_features=64)
    rng = np.random.default_rng(9)
    per = 6000
    c = abi.Context(0, max_points=len(kf.points) + 1000)
    c.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)
    conv = []
    for s in range(0, len(kf.points), per):
        scan = kf.points[s:s + per]
        rec, tb = rawfmt.livox_records(scan, WEEK, rng=rng, bad_fraction=0.03)
        raw = rawfmt.as_bytes(rec)
        want, _, _ = oracle.convert_raw(rawfmt.fmt_livox(), raw, len(rec), tb, 0.0, True, 6, 1.0, 200.0)
        m, _, _ = c.append_raw(rawfmt.fmt_livox(), raw, len(rec), timebase_ns=tb, scan_line=6)
        assert m == len(want)
        conv.append(want)
    pts = np.concatenate(conv)
    got = c.keyframe(kf.stamp, kf.ins, kf.queries)
    c2 = abi.Context(0)
    c2.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)
    c2.append_scan(pts)
    ref = c2.keyframe(kf.stamp, kf.ins, kf.queries)
    assert got["m"] == ref["m"] and got["n_projected"] == ref["n_projected"]
    assert np.array_equal(got["type"], ref["type"]) and got["xyz_depth"].tobytes() == ref["xyz_depth"].tobytes()
    orc = oracle.keyframe(pts, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries)
    assert got["m"] == orc["m"] and np.array_equal(got["type"], orc["type"])
    c.close()
    c2.close()


@pytest.mark.gpu
def test_append_raw_wraps_the_ring(oracle, built_lib):
    """Three scans 20 s apart in a ring that holds two and a half: the third is scattered across the ring's end, the
    eviction rule (le_vins.cc:156-159) has dropped the first two; the deskewed window equals the one from converted points."""
    from levins import abi

    kf = synth.make_keyframe(2, pts_per_scan=800, n_features=8)
    rng = np.random.default_rng(4)
    win = kf.points[(kf.points["time"] > kf.stamp - 0.45) & (kf.points["time"] <= kf.stamp)][:4000]
    c = abi.Context(0, max_points=10_000)
    c.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)
    last = None
    for shift in (-40.0, -20.0, 0.0):
        scan = win.copy()
        scan["time"] += shift
        rec, tb = rawfmt.livox_records(scan, WEEK, rng=rng, bad_fraction=0.02)
        raw = rawfmt.as_bytes(rec, 19)
        m, _, _ = c.append_raw(rawfmt.fmt_livox_wire(), raw, len(rec), timebase_ns=tb, scan_line=6)
        last, _, _ = oracle.convert_raw(rawfmt.fmt_livox_wire(), raw, len(rec), tb, 0.0, True, 6, 1.0, 200.0)
        assert m == len(last)
    size, t_front, t_back = c.buffer_info()
    assert size == len(last) and t_front == last["time"][0] and t_back == last["time"][-1]
    st, n, cloud = c.stack_deskew(kf.stamp, kf.ins)
    c2 = abi.Context(0)
    c2.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)
    c2.append_scan(last)
    st2, n2, cloud2 = c2.stack_deskew(kf.stamp, kf.ins)
    assert n == n2 > 0 and cloud.tobytes() == cloud2.tobytes()
    c.close()
    c2.close()


@pytest.mark.gpu
def test_append_raw_without_read_back_then_points(ctx, oracle):
    """levins_append_raw with no output pointers is asynchronous: the survivor count stays on the device and the next
    levins_append_scan has to learn the ring's tail from there.  Mixed sequence raw, raw, points, raw; the keyframe must
    equal the one computed from the oracle-converted points."""
    kf = synth.make_keyframe(2, pts_per_scan=5000, n_features=48)
    ctx.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)
    rng = np.random.default_rng(21)
    conv = []
    for i, s in enumerate(range(0, len(kf.points), 5000)):
        scan = kf.points[s:s + 5000]
        rec, tb = rawfmt.livox_records(scan, WEEK, rng=rng, bad_fraction=0.05)
        raw = rawfmt.as_bytes(rec)
        want, _, _ = oracle.convert_raw(rawfmt.fmt_livox(), raw, len(rec), tb, 0.0, True, 6, 1.0, 200.0)
        if i % 3 == 2:
            ctx.append_scan(want)  # PointXYZIT records from the host converter
        else:
            assert ctx.append_raw(rawfmt.fmt_livox(), raw, len(rec), timebase_ns=tb, scan_line=6, want_result=False) is None
        conv.append(want)
    pts = np.concatenate(conv)
    assert ctx.buffer_info()[0] == len(pts)
    got = ctx.keyframe(kf.stamp, kf.ins, kf.queries)
    ref = oracle.keyframe(pts, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries)
    assert got["n_projected"] == ref["n_projected"] and got["m"] == ref["m"]
    assert np.array_equal(got["type"], ref["type"]) and got["xyz_depth"].tobytes() == ref["xyz_depth"].tobytes()
