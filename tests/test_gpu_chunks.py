"""GPU parity for the paths of k_deskew_bin that ordinary data rarely takes ("chunks and staged entries"):
points whose INS interval is neither of the two staged with their chunk (dense INS, unsorted times), intervals whose
rotation is too large for the straight-line path, chunks that start before / end after the INS window, chunk sizes
around the 128-point boundary, and the queue of undecided cell ids."""
import numpy as np
import pytest

from levins import abi, synth

pytestmark = pytest.mark.gpu


def xyz(a):
    return np.stack([a["x"], a["y"], a["z"]], axis=-1)


def _check_stage1(ctx, oracle, kf, points, ins, min_identical=0.999):
    ctx.buffer_clear()
    ctx.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)
    ctx.append_scan(points)
    ref = oracle.stack_deskew(points, kf.stamp, ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5, details=True)
    st, n, cloud = ctx.stack_deskew(kf.stamp, ins)
    assert n == ref["n"]
    kept = ref["kept"].astype(bool)
    a, b = xyz(cloud), xyz(ref["cloud"])
    assert np.array_equal(np.any(a != 0, axis=1), kept)
    a, b = a[kept], b[kept]
    err = np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.linalg.norm(b.astype(np.float64), axis=1)
    assert err.max() <= 1e-4
    assert np.mean(np.all(a == b, axis=1)) >= min_identical
    # and the fused keyframe (deskew + binning in one kernel) against the oracle's keyframe
    ctx.buffer_clear()
    ctx.append_scan(points)
    got = ctx.keyframe(kf.stamp, ins, kf.queries)
    want = oracle.keyframe(points, kf.stamp, ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries)
    assert got["m"] == want["m"] and np.array_equal(got["type"], want["type"])
    return ref


def _spin_ins(ins, rate_rad_s, axis=(0.3, -0.5, 0.8)):
    """The same positions, orientation spinning at `rate_rad_s` about a fixed axis."""
    out = ins.copy()
    ax = np.asarray(axis, np.float64)
    ax /= np.linalg.norm(ax)
    ang = rate_rad_s * (ins["time"] - ins["time"][0])
    out["q"][:, :3] = np.sin(0.5 * ang)[:, None] * ax[None, :]  # Eigen coefficient order x, y, z, w
    out["q"][:, 3] = np.cos(0.5 * ang)
    return out


@pytest.mark.parametrize("rate", [2.0, 4.0, 30.0, 80.0])
def test_large_rotation_per_interval(ctx, oracle, rate):
    """3.1 rad/s is where theta^2 of a 5 ms interval crosses 2^-12: below it the straight-line path, above it the
    generic one (tier-2 polynomials, sincos beyond 0.25 rad)."""
    kf = synth.make_keyframe(2, pts_per_scan=6000, n_features=64)
    _check_stage1(ctx, oracle, kf, kf.points, _spin_ins(kf.ins, rate))


def test_dense_ins_more_than_two_intervals_per_chunk(ctx, oracle):
    """INS at 4 kHz: a 128-point chunk of the 240 kpts/s Avia spans 2-3 intervals, so some points find neither staged entry."""
    kf = synth.make_keyframe(2, pts_per_scan=6000, n_features=64)
    shp = synth.shape(2)
    ins = synth.make_ins(shp, kf.stamp - 0.6, kf.stamp + 0.1, rate=4000.0)  # 2 800 samples (config.max_ins is 4 096)
    ref = _check_stage1(ctx, oracle, kf, kf.points, ins)
    assert len(np.unique(ref["ins_index"])) > 1000


def test_sparse_ins(ctx, oracle):
    kf = synth.make_keyframe(2, pts_per_scan=6000, n_features=64)
    shp = synth.shape(2)
    for rate in (10.0, 3.0):  # 3 Hz: two or three samples over the 0.5 s stack
        ins = synth.make_ins(shp, kf.ins["time"][0], kf.ins["time"][-1], rate=rate)
        _check_stage1(ctx, oracle, kf, kf.points, ins)


def test_locally_unsorted_point_times(ctx, oracle):
    """Times shuffled inside 512-point blocks (away from the window ends, whose search needs order): the staged entries
    are only hints, every point is checked against its own time."""
    kf = synth.make_keyframe(2, pts_per_scan=6000, n_features=64)
    pts = kf.points.copy()
    rng = np.random.default_rng(11)
    inside = np.nonzero((pts["time"] > kf.stamp - 0.45) & (pts["time"] < kf.stamp - 0.05))[0]
    lo, hi = inside[0], inside[-1]
    for b in range(lo, hi - 512, 512):
        pts["time"][b:b + 512] = rng.permutation(pts["time"][b:b + 512])
    _check_stage1(ctx, oracle, kf, pts, kf.ins)


def test_window_edges_inside_chunks(ctx, oracle):
    """INS windows that begin and end in the middle of the stack at arbitrary points: chunks start in the fallback and
    move into the window, and leave it again (the fallback copy behind the last interval)."""
    kf = synth.make_keyframe(2, pts_per_scan=6000, n_features=64)
    n = len(kf.ins)
    for a, b in ((n // 3 + 1, 2 * n // 3 + 2), (n // 2, n // 2 + 2), (n // 2, n // 2 + 1), (0, n // 2 + 7)):
        ref = _check_stage1(ctx, oracle, kf, kf.points, kf.ins[a:b])
        assert (ref["ins_index"] == 0).any()


@pytest.mark.parametrize("n_pts", [1, 5, 127, 128, 129, 255, 257, 4097])
def test_chunk_boundaries(ctx, oracle, n_pts):
    kf = synth.make_keyframe(2, pts_per_scan=6000, n_features=16)
    inside = np.nonzero((kf.points["time"] > kf.stamp - 0.4) & (kf.points["time"] <= kf.stamp))[0]
    pts = kf.points[inside[0]:inside[0] + n_pts]
    ctx.buffer_clear()
    ctx.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)
    ctx.append_scan(pts)
    ref = oracle.stack_deskew(pts, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5, details=True)
    st, n, cloud = ctx.stack_deskew(kf.stamp, kf.ins)
    assert n == ref["n"] == n_pts
    assert xyz(cloud).tobytes() == xyz(ref["cloud"]).tobytes()


def test_many_undecided_cells_in_a_batch(oracle, built_lib):
    """A batch whose points sit on cell boundaries (camera = LiDAR = body, static): almost every cell id goes through the
    queue of undecided ids and its exact path; M and the winners must equal the oracle's range image."""
    rng = np.random.default_rng(3)
    ang = 0.3 * np.pi / 180.0
    n_kf, n = 3, 20_000
    c = abi.Context(0)
    eye = synth.pose_to_abi(np.eye(3), np.zeros(3))
    c.set_extrinsics(eye, eye, 0.0)
    pts_all, ins_all, q_all = [], [], []
    t0 = 1000.0
    for k in range(n_kf):
        row = rng.integers(0, 300, n)
        col = rng.integers(0, 300, n)
        # angles exactly on (or a few ulps off) the rounding boundaries k + 0.5
        el = (row - 150 + 0.5) * ang + rng.integers(-3, 4, n) * 1e-7
        az = (150 - col - 0.5) * ang + rng.integers(-3, 4, n) * 1e-7
        rngs = rng.uniform(2.0, 50.0, n)
        p = np.zeros(n, synth.XYZIT)
        p["x"] = (rngs * np.cos(el) * np.sin(az)).astype(np.float32)
        p["y"] = (rngs * np.sin(el)).astype(np.float32)
        p["z"] = (rngs * np.cos(el) * np.cos(az)).astype(np.float32)
        p["time"] = t0 + k + np.sort(rng.uniform(0.6, 1.0, n))
        ins = np.zeros(300, synth.INS)
        ins["time"] = ins["state_time"] = t0 + k + np.arange(300) / 200.0
        ins["q"][:, 3] = 1.0
        pts_all.append(p)
        ins_all.append(ins)
        q_all.append(np.array([[0, 0, 1.0]], np.float32))
    descs = np.zeros(n_kf, abi.KF_DESC)
    descs["pt_begin"], descs["pt_count"] = np.arange(n_kf) * n, n
    descs["ins_begin"], descs["ins_count"] = np.arange(n_kf) * 300, 300
    descs["q_begin"], descs["q_count"] = np.arange(n_kf), 1
    descs["stamp"] = t0 + np.arange(n_kf) + 1.0
    c.batch_upload(np.concatenate(pts_all), np.concatenate(ins_all), np.concatenate(q_all), descs)
    c.batch_run()
    out = c.batch_download()
    for k in range(n_kf):
        ref1 = oracle.stack_deskew(pts_all[k], t0 + k + 1.0, ins_all[k], eye, eye, 0.0, 0.5, details=True)
        ref = oracle.range_image(ref1["cloud"][: ref1["n"]])
        assert out["n_projected"][k] == ref1["n"] and out["m"][k] == ref["m"], (k, out["m"][k], ref["m"])
    c.close()


def test_ragged_batch_wave_sizes_and_keyframe_order(oracle, built_lib):
    """Keyframes of different sizes (one empty, one smaller than a chunk), different feature counts: every wave size and
    every keyframe order gives the same per-keyframe results, equal to the oracle's."""
    shp = synth.Shape(**{**synth.shape(2).__dict__, "pts_per_scan": 3000})
    b = synth.make_batch(2, 7, n_unique_scans=8, shp=shp)
    F = b.queries.shape[1]
    nq = [F, 5, 0, F, 17, F, 1]
    pt_begin, pt_count = b.pt_begin.copy(), b.pt_count.copy()
    pt_count[2] = 0            # empty keyframe
    pt_count[4] = 100          # fewer points than one chunk (they are the oldest of its window)
    pt_count[5] -= 777         # not a multiple of anything
    q_begin = np.concatenate([[0], np.cumsum(nq)[:-1]])
    queries = np.concatenate([b.queries[k][:nq[k]] for k in range(7)])

    def run(order, wave):
        d = np.zeros(7, abi.KF_DESC)
        d["pt_begin"], d["pt_count"] = pt_begin[order], pt_count[order]
        d["ins_begin"], d["ins_count"] = b.ins_begin[order], b.ins_count[order]
        d["q_begin"], d["q_count"] = q_begin[order], np.asarray(nq)[order]
        d["stamp"] = b.stamps[order]
        c = abi.Context(0)
        c.set_extrinsics(shp.rig.pose_b_c, shp.rig.pose_c_l, shp.rig.td_b_l)
        c.batch_set_wave(wave)
        c.batch_upload(b.points, b.ins, queries, d)
        c.batch_run()
        out = c.batch_download()
        c.close()
        return out

    ident = np.arange(7)
    base = run(ident, 256)
    for wave in (1, 3):
        o = run(ident, wave)
        assert all(np.array_equal(base[k], o[k]) for k in base), wave
    perm = np.array([3, 6, 0, 2, 5, 1, 4])
    o = run(perm, 2)
    assert np.array_equal(o["type"], base["type"]) and np.array_equal(o["xyz_depth"], base["xyz_depth"])  # results live at q_begin
    assert np.array_equal(o["n_projected"], base["n_projected"][perm]) and np.array_equal(o["m"], base["m"][perm])
    for k in range(7):
        pts = b.points[pt_begin[k]: pt_begin[k] + pt_count[k]]
        ins = b.ins[b.ins_begin[k]: b.ins_begin[k] + b.ins_count[k]]
        q = queries[q_begin[k]: q_begin[k] + nq[k]]
        if len(pts) == 0:
            assert base["n_projected"][k] == 0 and base["m"][k] == 0 and np.all(base["type"][q_begin[k]: q_begin[k] + nq[k]] == -1)
            continue
        ref = oracle.keyframe(pts, b.stamps[k], ins, shp.rig.pose_b_c, shp.rig.pose_c_l, shp.rig.td_b_l,
                              q if len(q) else np.zeros((0, 3), np.float32))
        assert base["n_projected"][k] == ref["n_projected"] and base["m"][k] == ref["m"], k
        assert np.array_equal(base["type"][q_begin[k]: q_begin[k] + nq[k]], ref["type"])
