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
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
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
    assert int(summary["single_type"]) == int(ref3["type"][0])
    single = np.array(summary["single"].split(), dtype=np.float32)
    assert np.allclose(single, ref3["xyz_depth"][0], rtol=1e-6)
