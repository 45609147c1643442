"""GPU edge cases through the C ABI (SURVEY 8a note 7): empty inputs, everything outside the FOV, fewer than 5 cells,
timestamps outside the INS window, irregular INS spacing, ring wrap-around and the 11 s eviction rule, a streaming
sequence of keyframes with persistent pops, capacity and state errors."""
import numpy as np
import pytest

from levins import abi, synth

pytestmark = pytest.mark.gpu


def xyz(a):
    return np.stack([a["x"], a["y"], a["z"]], axis=-1)


def _setup(ctx, kf):
    ctx.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)


def test_empty_ring_and_empty_window(ctx, oracle):
    kf = synth.make_keyframe(2, pts_per_scan=1000, n_features=16)
    _setup(ctx, kf)
    # nothing appended yet: the reference never calls the stacker with an empty buffer (le_vins.cc:378); EMPTY here
    r = ctx.keyframe(kf.stamp, kf.ins, kf.queries)
    assert r["status"] == abi.EMPTY and np.all(r["type"] == -1) and np.array_equal(r["xyz_depth"][:, :3], kf.queries)
    st, n, cloud = ctx.stack_deskew(kf.stamp, kf.ins, capacity=0)
    assert st == abi.EMPTY and n == 0
    # every point older than the window: all popped, "No valid point cloud" (le_vins.cc:426-429)
    ctx.append_scan(kf.points)
    st, n, cloud = ctx.stack_deskew(kf.stamp + 50.0, kf.ins)
    assert st == abi.EMPTY and n == 0 and ctx.buffer_info()[0] == 0
    # every point newer than the stamp: nothing projected, nothing popped
    ctx.append_scan(kf.points)
    st, n, cloud = ctx.stack_deskew(kf.points["time"][0] - 1.0, kf.ins)
    assert st == abi.EMPTY and n == 0 and ctx.buffer_info()[0] == len(kf.points)


def test_everything_outside_the_fov_and_tiny_images(ctx, oracle):
    behind = np.zeros(5000, synth.XYZI)
    rng = np.random.default_rng(0)
    behind["x"], behind["y"] = rng.normal(0, 5, 5000), rng.normal(0, 5, 5000)
    behind["z"] = -np.abs(rng.normal(10, 3, 5000))
    behind["w"] = 1.0
    st, m, unit, winner = ctx.build_range_image(behind)
    assert st == abi.EMPTY and m == 0 and np.all(winner == -1)  # tracking.cc:356-358
    empty = np.zeros(0, synth.XYZI)
    st, m, unit, winner = ctx.build_range_image(empty)
    assert st == abi.EMPTY and m == 0
    with pytest.raises(abi.LevinsError):  # no image -> associate is a state error, not a silent NONE
        ctx.associate(np.array([[0, 0, 1.0]], np.float32))
    # default (zero) points: 0/0 = NaN passes the FOV gate, col = -150 drops them
    zeros = np.zeros(100, synth.XYZI)
    zeros["w"] = 1.0
    assert ctx.build_range_image(zeros)[1] == 0
    # M < 5: the reference reads k_sqr_distances[4] out of bounds; the build defines NONE
    few = np.zeros(3, synth.XYZI)
    few["z"] = [5.0, 5.0, 5.0]
    few["x"] = [0.0, 0.03, -0.03]
    st, m, unit, winner = ctx.build_range_image(few)
    ref = oracle.range_image(few)
    assert m == 3 == ref["m"] and unit.tobytes() == ref["unit_sphere"].tobytes()
    got = ctx.associate(np.array([[0, 0, 1.0]], np.float32))
    assert got["type"][0] == -1 and got["flags"][0] == 1 and np.array_equal(got["knn_idx"][0][3:], [-1, -1])
    assert np.array_equal(got["knn_idx"][0][:3], oracle.brute_knn(ref["unit_sphere"], [0, 0, 1.0])[1][:3])


def test_timestamps_outside_the_ins_window(ctx, oracle):
    kf = synth.make_keyframe(2, pts_per_scan=4000, n_features=64)
    _setup(ctx, kf)
    for ins in (kf.ins[: int(len(kf.ins) * 0.85)],  # window ends inside the stack: newest-state fallback (misc.cc:74-77)
                kf.ins[int(len(kf.ins) * 0.8):],     # window starts inside the stack: time < front also falls back
                kf.ins[-1:]):                        # a single sample: always the fallback
        ctx.buffer_clear()
        ctx.append_scan(kf.points)
        ref = oracle.stack_deskew(kf.points, kf.stamp, ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5, details=True)
        st, n, cloud = ctx.stack_deskew(kf.stamp, ins)
        assert n == ref["n"] and (ref["ins_index"] == 0).any()
        kept = ref["kept"].astype(bool)
        a, b = xyz(cloud)[kept], xyz(ref["cloud"])[kept]
        err = np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.linalg.norm(b.astype(np.float64), axis=1)
        assert err.max() <= 1e-4 and np.mean(np.all(a == b, axis=1)) > 0.999


def test_irregular_ins_spacing(ctx, oracle):
    """Dropped and jittered INS samples defeat the uniform-spacing guess; the bisection fallback must agree."""
    kf = synth.make_keyframe(2, pts_per_scan=4000)
    _setup(ctx, kf)
    rng = np.random.default_rng(5)
    keep = np.ones(len(kf.ins), bool)
    keep[rng.choice(np.arange(1, len(kf.ins) - 1), size=len(kf.ins) // 3, replace=False)] = False
    keep[40:90] = False  # a 0.25 s gap
    ins = kf.ins[keep].copy()
    ref = oracle.stack_deskew(kf.points, kf.stamp, ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5, details=True)
    ctx.append_scan(kf.points)
    st, n, cloud = ctx.stack_deskew(kf.stamp, ins)
    kept = ref["kept"].astype(bool)
    a, b = xyz(cloud)[kept], xyz(ref["cloud"])[kept]
    assert n == ref["n"] and np.mean(np.all(a == b, axis=1)) > 0.999
    assert np.max(np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.linalg.norm(b.astype(np.float64), axis=1)) <= 1e-4


def test_ring_eviction_and_wraparound(oracle, built_lib):
    """addNewPointCloud's 11 s rule (le_vins.cc:156-159) on a ring small enough to wrap several times."""
    c = abi.Context(0, max_points=40_000)
    ring = oracle.Ring()
    rng = np.random.default_rng(7)
    t = 1000.0
    for scan in range(80):
        n = int(rng.integers(500, 3000))
        pts = np.zeros(n, synth.XYZIT)
        pts["time"] = t + np.sort(rng.random(n)) * 1.0
        pts["x"] = scan
        t += 1.0
        c.append_scan(pts)
        ring.append(pts)
        if scan % 9 == 0:
            size, t_front, t_back = c.buffer_info()
            data = ring.data()
            assert size == len(ring) and t_front == data["time"][0] and t_back == data["time"][-1]
    assert c.buffer_info()[0] == len(ring) < 40_000
    with pytest.raises(abi.LevinsError) as e:  # more live points than the ring holds
        big = np.zeros(45_000, synth.XYZIT)
        big["time"] = t + np.linspace(0, 0.5, 45_000)
        c.append_scan(big)
    assert e.value.status == abi.ERR_CAPACITY
    c.close()


def test_streaming_sequence_with_wrapping_ring(oracle, built_lib):
    """Several keyframes in a row on one context: pops persist (the deque is shared state, le_vins.cc:413-424), scans keep
    arriving, the ring wraps; every keyframe equals the oracle run on the oracle's own ring."""
    shp = synth.Shape(**{**synth.shape(2).__dict__, "pts_per_scan": 5000})
    b = synth.make_batch(2, 8, n_unique_scans=11, shp=shp)
    c = abi.Context(0, max_points=80_000)  # 16 scans: wraps several times during the sequence
    c.set_extrinsics(shp.rig.pose_b_c, shp.rig.pose_c_l, shp.rig.td_b_l)
    ring = oracle.Ring()
    per = shp.pts_per_scan
    scan = 0
    n_scans = len(b.points) // per
    for k in range(b.n_keyframes):
        stamp = float(b.stamps[k])
        ins = b.ins[b.ins_begin[k]: b.ins_begin[k] + b.ins_count[k]]
        while scan < n_scans and b.points["time"][scan * per] <= stamp:
            s = b.points[scan * per:(scan + 1) * per]
            c.append_scan(s)
            ring.append(s)
            scan += 1
        buf = ring.data()
        ref = oracle.keyframe(buf, stamp, ins, shp.rig.pose_b_c, shp.rig.pose_c_l, shp.rig.td_b_l, b.queries[k], 0.5)
        ring.evict_before(stamp - 0.5)
        if k % 2 == 0:
            got = c.keyframe(stamp, ins, b.queries[k])
            assert got["n_projected"] == ref["n_projected"] and got["m"] == ref["m"]
            ty, xd = got["type"], got["xyz_depth"]
        else:  # stage-wise on the same context
            st, n, cloud = c.stack_deskew(stamp, ins)
            assert n == ref["n_projected"]
            st, m, unit, winner = c.build_range_image(None)
            assert m == ref["m"]
            a = c.associate(b.queries[k])
            ty, xd = a["type"], a["xyz_depth"]
        assert c.buffer_info()[0] == len(ring)
        assert np.array_equal(ty, ref["type"])
        both = ty == 1
        assert both.sum() > 20
        assert np.max(np.abs(xd[both, 3] - ref["xyz_depth"][both, 3]) / ref["xyz_depth"][both, 3]) <= 1e-4
    c.close()


def test_capacity_and_argument_errors(ctx):
    kf = synth.make_keyframe(2, pts_per_scan=500, n_features=8)
    with pytest.raises(abi.LevinsError) as e:  # extrinsics not set
        ctx.append_scan(kf.points)
        ctx.keyframe(kf.stamp, kf.ins, kf.queries)
    assert e.value.status == abi.ERR_STATE
    _setup(ctx, kf)
    with pytest.raises(abi.LevinsError) as e:
        ctx.keyframe(kf.stamp, kf.ins[:0], kf.queries)  # empty INS window: the reference would read back() of an empty deque
    assert e.value.status == abi.ERR_ARG
    small = abi.Context(0, max_features=4)
    small.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)
    small.append_scan(kf.points)
    with pytest.raises(abi.LevinsError) as e:
        small.keyframe(kf.stamp, kf.ins, kf.queries)
    assert e.value.status == abi.ERR_CAPACITY
    small.close()


def test_full_size_properties_config3(built_lib):
    """BASELINE configs[2] at full size (2.6 M points / keyframe, td_b_l = -0.48 s): size-independent properties.
    Idempotence (re-running a resident batch reproduces every output bit), permutation (keyframe order does not matter),
    and monotonicity (appending points can only add cells)."""
    b = synth.make_batch(3, 3, n_unique_scans=12, ins_span=3.5)
    c = abi.Context(0, stack_interval=b.shape.stack_interval)
    c.set_extrinsics(b.shape.rig.pose_b_c, b.shape.rig.pose_c_l, b.shape.rig.td_b_l)
    d = abi.batch_descs(b)
    c.batch_upload(b.points, b.ins, b.queries.reshape(-1, 3), d)
    c.batch_run()
    o1 = c.batch_download()
    c.batch_run()
    o2 = c.batch_download()
    assert all(np.array_equal(o1[k], o2[k]) for k in o1)
    assert o1["n_projected"].min() > 2_500_000  # 20 scans x 131 072 points
    assert o1["m"].min() > 10_000 and (o1["type"] == 1).mean() > 0.2
    perm = np.array([2, 0, 1])
    c.batch_upload(b.points, b.ins, b.queries.reshape(-1, 3), d[perm])
    c.batch_run()
    o3 = c.batch_download()
    f = b.queries.shape[1]
    for i, k in enumerate(perm):  # results follow the descriptors' q_begin, i.e. stay with the keyframe
        assert o3["n_projected"][i] == o1["n_projected"][k] and o3["m"][i] == o1["m"][k]
    assert np.array_equal(o3["type"], o1["type"]) and np.array_equal(o3["xyz_depth"], o1["xyz_depth"])
    # a keyframe restricted to the first half of its buffer has a subset of the cells
    half = d[:1].copy()
    half["pt_count"] = half["pt_count"] // 2
    c.batch_upload(b.points, b.ins, b.queries.reshape(-1, 3), half)
    c.batch_run()
    o4 = c.batch_download()
    assert o4["m"][0] <= o1["m"][0] and o4["n_projected"][0] <= o1["n_projected"][0]
    c.close()


def test_scans_from_page_locked_buffers(ctx, oracle):
    """levins_alloc_host / levins_free_host: the same keyframe whether the scans come from pageable or page-locked memory."""
    kf = synth.make_keyframe(2, pts_per_scan=4000, n_features=64)
    _setup(ctx, kf)
    ctx.append_scan(kf.points)
    want = ctx.keyframe(kf.stamp, kf.ins, kf.queries)
    hb = abi.HostBuffer(len(kf.points))
    hb.array[:] = kf.points
    ctx.buffer_clear()
    half = len(kf.points) // 2
    ctx.append_scan(hb.array[:half])
    ctx.append_scan(hb.array[half:])
    got = ctx.keyframe(kf.stamp, kf.ins, kf.queries)
    assert got["m"] == want["m"] and got["n_projected"] == want["n_projected"]
    assert np.array_equal(got["type"], want["type"]) and got["xyz_depth"].tobytes() == want["xyz_depth"].tobytes()
    ref = oracle.keyframe(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries)
    assert np.array_equal(got["type"], ref["type"]) and got["m"] == ref["m"]
    hb.close()


def test_more_than_254_waves(oracle, built_lib):
    """z-buffer slots are recycled by generation and cleared for real once every 254 uses: a batch with more waves than
    that (wave = 1, some keyframes without features, so that no feature kernel separates one wave's cell count from the
    next wave's clear) must still report every keyframe's M."""
    shp = synth.Shape(**{**synth.shape(2).__dict__, "pts_per_scan": 600, "n_features": 4})
    b = synth.make_batch(2, 300, n_unique_scans=9, shp=shp)
    descs = abi.batch_descs(b)
    descs["q_count"][::3] = 0  # every third keyframe has no features
    c = abi.Context(0)
    c.set_extrinsics(b.shape.rig.pose_b_c, b.shape.rig.pose_c_l, b.shape.rig.td_b_l)
    c.batch_set_wave(1)
    c.batch_upload(b.points, b.ins, b.queries.reshape(-1, 3), descs)
    want_m = None
    for rep in range(2):  # the second run starts in the middle of a generation cycle
        c.batch_run()
        out = c.batch_download()
        if want_m is None:
            want_m = np.zeros(b.n_keyframes, np.uint32)
            for k in range(b.n_keyframes):
                kf = synth.batch_keyframe(b, k)
                r1 = oracle.stack_deskew(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5)
                want_m[k] = oracle.range_image(r1["cloud"])["m"]
        assert np.array_equal(out["m"], want_m), np.nonzero(out["m"] != want_m)[0][:10]
    c.close()


def _edge_band_scenario(seed, n=6000):
    """Points aimed at the border of the 90-degree image cone (|x| / z or |y| / z = 1 +- 1e-7 .. 1e-1), every one at its own
    border cell, seen from a slowly moving rig: the certified frustum cull of k_deskew_bin must drop exactly the points the
    reference's FOV gate (tracking.cc:311-314) drops."""
    rng = np.random.default_rng(seed)
    t0 = float(rng.uniform(1e3, 6e5))
    stamp = t0 + 2.0
    times = np.sort(rng.uniform(stamp - 0.5, stamp, n))
    side = rng.integers(0, 4, n)
    along = np.tan(np.deg2rad(rng.uniform(-44.9, 44.9, n)))
    off = 1.0 + rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-7, -1, n)
    x = np.where(side == 0, off, np.where(side == 1, -off, along))
    y = np.where(side == 2, off, np.where(side == 3, -off, along))
    d = np.stack([x, y, np.ones(n)], axis=1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rng.uniform(2.0, 150.0, n)
    R_lc = np.eye(3)
    pts = np.zeros(n, synth.XYZIT)
    p = d * r[:, None]
    pts["x"], pts["y"], pts["z"] = p[:, 0].astype(np.float32), p[:, 1].astype(np.float32), p[:, 2].astype(np.float32)
    pts["time"] = times
    rate = float(rng.choice([50.0, 200.0]))
    tt = np.arange(stamp - 0.7, stamp + 0.1, 1.0 / rate)
    ins = np.zeros(len(tt), synth.INS)
    ins["time"] = tt
    ins["state_time"] = tt
    w = rng.normal(size=3) * rng.choice([0.0, 0.02, 0.3])
    v = rng.normal(size=3) * rng.choice([0.0, 0.1, 3.0])
    ang = np.linalg.norm(w) * (tt - stamp)
    ax = w / max(np.linalg.norm(w), 1e-12)
    ins["q"][:, :3] = np.sin(0.5 * ang)[:, None] * ax
    ins["q"][:, 3] = np.cos(0.5 * ang)
    ins["p"] = v[None, :] * (tt - stamp)[:, None]
    eye = synth.pose_to_abi(R_lc, np.zeros(3))
    return dict(pts=pts, ins=ins, stamp=stamp, pose_b_c=eye, pose_c_l=eye)


@pytest.mark.parametrize("seed", range(6))
def test_frustum_cull_is_exact_on_the_fov_border(oracle, built_lib, seed):
    s = _edge_band_scenario(100 + seed)
    c = abi.Context(0)
    c.set_extrinsics(s["pose_b_c"], s["pose_c_l"], 0.0)
    c.append_scan(s["pts"])
    q = np.array([[0.0, 0.0, 1.0]], np.float32)
    got = c.keyframe(s["stamp"], s["ins"], q)
    ref1 = oracle.stack_deskew(s["pts"], s["stamp"], s["ins"], s["pose_b_c"], s["pose_c_l"], 0.0, 0.5)
    ref2 = oracle.range_image(ref1["cloud"])
    assert got["n_projected"] == ref1["n"]
    assert 200 < ref2["m"] < len(s["pts"])  # both sides of the border are populated
    assert got["m"] == ref2["m"], (got["m"], ref2["m"])
    c.close()


def test_results_do_not_depend_on_the_cull_policy(oracle, built_lib):
    """The library switches between the deskew instantiations with and without the frustum cull from what the cull
    statistics say (all points inside the image cone: off after the first read-back, a probe every 64th launch).  Seventy
    runs of one batch walk through on -> off -> probe -> off; every run must give the same bytes."""
    shp = synth.Shape(**{**synth.shape(2).__dict__, "pts_per_scan": 4000, "n_features": 40})
    b = synth.make_batch(2, 5, n_unique_scans=9, shp=shp)
    c = abi.Context(0)
    c.set_extrinsics(b.shape.rig.pose_b_c, b.shape.rig.pose_c_l, b.shape.rig.td_b_l)
    c.batch_upload(b.points, b.ins, b.queries.reshape(-1, 3), abi.batch_descs(b))
    first = None
    for run in range(70):
        c.batch_run()
        out = c.batch_download()
        if first is None:
            first = out
            kf = synth.batch_keyframe(b, 2)
            ref = oracle.keyframe(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries, 0.5)
            assert out["m"][2] == ref["m"] and np.array_equal(out["type"][80:120], ref["type"])
        else:
            assert all(out[k].tobytes() == first[k].tobytes() for k in out), run
    c.close()
