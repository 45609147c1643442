"""The C++ host mirror of the reference's class / function surface (le-vins_b200/host) driven like the estimator
drives it: addNewPointCloud per scan -> pointcloudStackForVisualTracking -> depthAssociationProjection ->
depthAssociationReference, compared with the oracle."""
import importlib.util
import os
import subprocess

import numpy as np
import pytest

from levins import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    spec = importlib.util.spec_from_file_location("levins_host_build", os.path.join(ROOT, "le-vins_b200", "host", "build_host.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def _plane_patches(rng):
    """patches of 5 nearly coplanar points + noise so that planeEstimation sees both outcomes"""
    pts = np.zeros(200, synth.XYZI)
    base = rng.normal(0, 8, (40, 3)) + [0, 0, 15]
    nrm = rng.normal(0, 1, (40, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    for g in range(40):
        off = rng.normal(0, 0.4, (5, 3))
        if g % 3:  # two thirds of the patches are planar (1 cm roughness), the rest scatter 0.4 m off the plane
            off -= np.outer(off @ nrm[g], nrm[g]) * 0.975
        p = (base[g] + off).astype(np.float32)
        pts["x"][5 * g:5 * g + 5], pts["y"][5 * g:5 * g + 5], pts["z"][5 * g:5 * g + 5] = p[:, 0], p[:, 1], p[:, 2]
    pts["w"] = 1.0
    pts["intensity"] = rng.uniform(0, 100, 200).astype(np.float32)
    return pts


def _check_plane_outputs(oracle, pts, plane, stds, plane_batch):
    """planeEstimation bit for bit against the oracle, pointToPlaneStd against the oracle (1e-12) and against an independent
    numpy restatement of pointcloud.cc:111-190 with numpy.linalg.eigh (1e-6)"""
    import ctypes

    n_ok = 0
    for g in range(40):
        five = pts[5 * g:5 * g + 5]
        ok, n, ninv, dist = oracle.plane_estimation(five, 0.1)
        assert bool(plane[g, 0]) == ok and np.array_equal(plane[g, 1:4], n) and plane[g, 4] == ninv and int(plane[g, 5]) == len(dist)
        assert bool(plane_batch[g]) == ok
        n_ok += ok
        target = pts[(5 * g + 7) % len(pts)]
        s_or = oracle.point_to_plane_std(five, target, 0.03)
        assert np.isclose(stds[g, 0], s_or, rtol=1e-12)
        P = np.stack([five["x"], five["y"], five["z"]], -1).astype(np.float64)
        c = P.mean(0)
        cov = P.T @ P / 5 - np.outer(c, c)
        w, V = np.linalg.eigh(cov)
        nn = V[:, 0]
        pc = np.zeros((6, 6))
        for i in range(5):
            F = np.zeros((3, 3))
            for m in (1, 2):
                F[m] = (1 / 5) * (P[i] - c) / (w[0] - w[m]) @ (np.outer(V[:, m], nn) + np.outer(nn, V[:, m]))
            J = np.vstack([V @ F, np.eye(3) / 5])
            pc += J @ (np.eye(3) * 0.03**2) @ J.T
        jac = np.concatenate([np.array([target["x"], target["y"], target["z"]], np.float64) - c, -nn])
        ref = np.sqrt(jac @ pc @ jac + 0.03**2 * nn @ nn)
        assert np.isclose(s_or, ref, rtol=1e-6), (s_or, ref)
        assert stds[g, 1] == np.float32(oracle.lib().orc_point_square_distance(
            five[0:1].ctypes.data_as(ctypes.c_void_p), five[1:2].ctypes.data_as(ctypes.c_void_p)))
    assert 5 < n_ok < 35  # both planeEstimation outcomes exercised


def _dump(kf, d, frame_pose, n_scans):
    kf.points.tofile(os.path.join(d, "points.bin"))
    ins = np.concatenate([kf.ins["time"][:, None], kf.ins["state_time"][:, None], kf.ins["p"], kf.ins["q"]], axis=1)
    ins.astype(np.float64).tofile(os.path.join(d, "ins.bin"))
    kf.pixels.astype(np.float32).tofile(os.path.join(d, "pixels.bin"))
    meta = np.concatenate([[kf.stamp, kf.rig.td_b_l], kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.intr, frame_pose, [n_scans]])
    meta.astype(np.float64).tofile(os.path.join(d, "meta.bin"))


def test_host_mirror_builds_and_refuses_without_gpu(built_lib, tmp_path):
    exe = _build()
    assert os.path.exists(exe)
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the refusal path is checked on the CPU box")
    kf = synth.make_keyframe(2, pts_per_scan=300, n_features=8)
    _dump(kf, str(tmp_path), np.concatenate([np.eye(3).reshape(-1), np.zeros(3)]), 3)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_host_mirror_matches_oracle(oracle, built_lib, tmp_path):
    exe = _build()
    kf = synth.make_keyframe(2, pts_per_scan=6000, n_features=120)
    ok, frame_pose = oracle.get_pose_from_ins_window(kf.ins, kf.rig.pose_b_c, kf.stamp)
    n_scans = 8
    _dump(kf, str(tmp_path), frame_pose, n_scans)
    # the same scans as Livox driver records for LidarConverter::livoxPointCloudAppend (device-side conversion)
    from levins import rawfmt

    rng = np.random.default_rng(12)
    per = (len(kf.points) + n_scans - 1) // n_scans
    recs, meta, conv = [], [], []
    for b in range(0, len(kf.points), per):
        rec, tb = rawfmt.livox_records(kf.points[b:b + per], 2334, rng=rng, bad_fraction=0.03)
        recs.append(rec)
        meta += [tb, len(rec)]
        conv.append(oracle.convert_raw(rawfmt.fmt_livox(), rawfmt.as_bytes(rec), len(rec), tb, 0.0, True, 6, 1.0, 200.0)[0])
    np.concatenate(recs).tofile(tmp_path / "raw_livox.bin")
    np.array(meta, np.uint64).tofile(tmp_path / "raw_timebase.bin")
    # the three PointCloud2 drivers' records for LidarConverter::{velodyne,ouster,hesai}PointCloudAppend
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_raw_formats import _cases

    cases = {c[0]: c for c in _cases(seed=5, n=20_000)}
    drv_stamp = cases["velodyne"][4]
    np.array([drv_stamp, 1.0, 200.0, 1.0, 128], np.float64).tofile(tmp_path / "raw_meta.bin")
    for name in ("velodyne", "ouster", "hesai"):
        cases[name][2].tofile(tmp_path / f"raw_{name}.bin")
    patches = _plane_patches(np.random.default_rng(11))
    patches.tofile(tmp_path / "pts.bin")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    _check_plane_outputs(oracle, patches, np.fromfile(tmp_path / "plane.bin").reshape(-1, 6),
                         np.fromfile(tmp_path / "std.bin").reshape(-1, 2), np.fromfile(tmp_path / "plane_batch.bin"))
    drv = np.fromfile(tmp_path / "append_drivers.bin", np.float64).reshape(3, 4)
    for i, name in enumerate(("velodyne", "ouster", "hesai")):
        _, fmt, rec, tb, st, lines = cases[name]
        want, wst, wen = oracle.convert_raw(fmt, rawfmt.as_bytes(rec), len(rec), tb, st, True, 128, 1.0, 200.0)
        assert int(drv[i, 0]) == len(want) == int(drv[i, 1]) > 0 and drv[i, 2] == wst and drv[i, 3] == wen, name
    sraw = dict(line.split(" ", 1) for line in open(tmp_path / "summary_raw.txt").read().strip().split("\n"))
    pts_conv = np.concatenate(conv)
    ref_raw = oracle.stack_deskew(pts_conv, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5)
    assert int(sraw["appended"]) == len(pts_conv) and int(sraw["n"]) == ref_raw["n"] > 0
    assert np.fromfile(tmp_path / "cloud_raw.bin", dtype=synth.XYZI).tobytes() == ref_raw["cloud"].tobytes()
    summary = dict(line.split(" ", 1) for line in open(tmp_path / "summary.txt").read().strip().split("\n"))
    ref1 = oracle.stack_deskew(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, 0.5)
    ref2 = oracle.range_image(ref1["cloud"])
    ref3 = oracle.associate(ref2["unit_sphere"], kf.queries)
    assert int(summary["stacked"]) == 1 and int(summary["n"]) == ref1["n"] and int(summary["M"]) == ref2["m"]
    cloud = np.fromfile(tmp_path / "cloud.bin", dtype=synth.XYZI)
    assert cloud.tobytes() == ref1["cloud"].tobytes()
    assoc = np.fromfile(tmp_path / "assoc.bin", dtype=np.float64).reshape(-1, 6)
    # the driver's gate: type != NONE and isGoodDepth(depth, 1) (tracking.cc:431-435)
    good = (ref3["type"] != -1) & (ref3["xyz_depth"][:, 3] > 1.0) & (ref3["xyz_depth"][:, 3] < 200.0)
    assert int(summary["associated"]) == int(good.sum()) == len(assoc) and len(assoc) > 20
    assert np.array_equal(assoc[:, 0].astype(int), np.nonzero(good)[0])
    assert np.array_equal(assoc[:, 4].astype(np.float32), ref3["xyz_depth"][good, 3])
    R, t = frame_pose[:9].reshape(3, 3).T, frame_pose[9:]
    pw = ref3["xyz_depth"][good, :3].astype(np.float64) @ R.T + t  # cam2world (camera.cc:149-151)
    assert np.allclose(assoc[:, 1:4], pw, rtol=0, atol=1e-9)
    # skip branches of the two drivers (tracking.cc:401-404, 513-516): skipped features are not evaluated, the rest unchanged
    idx = np.arange(len(good))
    for name, skipped in (("assoc_ref_skip.bin", idx % 3 == 0), ("assoc_map_skip.bin", idx % 4 == 1)):
        o = np.fromfile(tmp_path / name, np.float64)
        lst = o[1:].reshape(-1, 3)
        want_idx = np.nonzero(good & ~skipped)[0]
        assert int(o[0]) == len(want_idx) == len(lst) and 0 < len(want_idx) < int(good.sum()), name
        assert np.array_equal(lst[:, 0].astype(int), want_idx)
        assert np.array_equal(lst[:, 1].astype(np.float32), ref3["xyz_depth"][want_idx, 3]) and np.all(lst[:, 2] == 1)
    assert int(summary["single_type"]) == int(ref3["type"][0])
    single = np.array(summary["single"].split(), dtype=np.float32)
    assert np.allclose(single, ref3["xyz_depth"][0], rtol=1e-6)


def test_host_mirror_cpu_functions_match_oracle(oracle, built_lib, tmp_path):
    """The part of the C++ mirror that needs no GPU: MISC pose lookup (host exports of the library), PointCloudCommon single-point statics,
    Camera -- bit-for-bit against the oracle (same operation order, -ffp-contract=off)."""
    _build()
    exe = os.path.join(ROOT, "le-vins_b200", "host", "test_host_cpu")
    kf = synth.make_keyframe(2, pts_per_scan=400, n_features=8)
    rig = kf.rig
    ins = np.concatenate([kf.ins["time"][:, None], kf.ins["state_time"][:, None], kf.ins["p"], kf.ins["q"]], axis=1)
    ins.astype(np.float64).tofile(tmp_path / "ins.bin")
    rng = np.random.default_rng(11)
    t = kf.ins["time"]
    times = np.concatenate([rng.uniform(t[0], t[-1], 60), [t[0] - 1.0, t[0], t[5], t[-1], t[-1] + 2.0]])
    times.tofile(tmp_path / "times.bin")
    np.concatenate([rig.pose_b_c, rig.pose_c_l, rig.intr]).tofile(tmp_path / "ext.bin")
    pts = _plane_patches(rng)
    pts.tofile(tmp_path / "pts.bin")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

    poses = np.fromfile(tmp_path / "poses.bin").reshape(-1, 14)
    pose_list = []
    for k, tm in enumerate(times):
        ok, pose = oracle.get_pose_from_ins_window(kf.ins, rig.pose_b_c, tm)
        assert int(poses[k, 0]) == oracle.get_ins_window_index(kf.ins, tm) and bool(poses[k, 1]) == ok
        assert np.array_equal(poses[k, 2:], pose), f"pose {k} differs"
        pose_list.append(pose)
    proj = np.fromfile(tmp_path / "proj.bin", dtype=synth.XYZI).reshape(-1, 3)
    for k in range(len(pts)):
        p0, p1 = pose_list[k % len(pose_list)], pose_list[(k * 7 + 3) % len(pose_list)]
        assert proj[k, 0].tobytes() == oracle.point_projection(rig.pose_c_l, pts[k]).tobytes()
        assert proj[k, 1].tobytes() == oracle.relative_projection(p0, p1, pts[k]).tobytes()
        assert proj[k, 2].tobytes() == oracle.relative_projection_to_camera(rig.pose_c_l, p0, p1, pts[k]).tobytes()
    cam = np.fromfile(tmp_path / "cam.bin").reshape(-1, 10)
    for k in range(len(pts)):
        u, v = np.float32(pts["x"][k] * np.float32(10.0) + np.float32(600.0)), np.float32(pts["y"][k] * np.float32(10.0) + np.float32(500.0))
        assert np.array_equal(cam[k, 0:3], oracle.pixel2cam(rig.intr, u, v))
        puc = oracle.pixel2unitcam(rig.intr, u, v)
        assert np.array_equal(cam[k, 3:6], puc)
        assert np.array_equal(cam[k, 6:9], oracle.cam2world(puc, pose_list[k % len(pose_list)]))
        assert bool(cam[k, 9]) == oracle.is_good_depth(pts["z"][k], 1.0)
