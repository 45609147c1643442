"""levins_config.deskew_precision = 1 (f32 pose algebra): north_star's floating-point bar is 1e-4 relative for the deskewed
coordinates and the depths, and >= 95 % agreement of the neighbour / validity decisions.  The default mode is bit-exact
(test_gpu_parity.py); this mode trades that for throughput and must stay far inside the bar."""
import numpy as np
import pytest

from levins import abi, synth

pytestmark = pytest.mark.gpu


def xyz(a):
    return np.stack([a["x"], a["y"], a["z"]], axis=-1)


@pytest.mark.parametrize("case", [dict(cfg=1, pts_per_scan=8000), dict(cfg=2, pts_per_scan=8000), dict(cfg=4)])
def test_f32_mode_against_the_oracle(oracle, built_lib, case):
    kf = synth.make_keyframe(**case)
    c = abi.Context(0, deskew_precision=1, stack_interval=kf.shape.stack_interval)
    c.set_extrinsics(kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l)
    c.append_scan(kf.points)
    si = c.cfg.stack_interval
    ref = oracle.stack_deskew(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, si, details=True)
    st, n, cloud = c.stack_deskew(kf.stamp, kf.ins)
    # window membership and range gate do not depend on the pose algebra: still bit-exact
    assert n == ref["n"]
    kept = ref["kept"].astype(bool)
    a, b = xyz(cloud), xyz(ref["cloud"])
    assert np.array_equal(np.any(a != 0, axis=1), kept)
    err = np.linalg.norm(a[kept].astype(np.float64) - b[kept], axis=1) / np.linalg.norm(b[kept].astype(np.float64), axis=1)
    print(f"f32 mode: max rel err {err.max():.2e}, identical {np.mean(np.all(a[kept] == b[kept], axis=1)):.3f}")
    assert err.max() <= 1e-5  # bar 1e-4
    # whole keyframe: validity decisions and depths
    c.buffer_clear()
    c.append_scan(kf.points)
    got = c.keyframe(kf.stamp, kf.ins, kf.queries)
    want = oracle.keyframe(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries, si)
    agree = np.mean(got["type"] == want["type"])
    both = (got["type"] == 1) & (want["type"] == 1)
    rel = np.abs(got["xyz_depth"][both, 3].astype(np.float64) - want["xyz_depth"][both, 3]) / np.abs(want["xyz_depth"][both, 3])
    print(f"f32 mode: validity agreement {agree:.4f}, depth within 1e-4: {np.mean(rel <= 1e-4):.4f}, M {got['m']} vs {want['m']}")
    assert agree >= 0.95
    assert abs(int(got["m"]) - int(want["m"])) <= max(3, want["m"] // 1000)
    # depths of features whose 5 neighbours are the same cells agree to the bar; a different neighbour set (a point that
    # changed cells) legitimately gives another plane, so the bound is on the bulk
    assert both.sum() == 0 or np.mean(rel <= 1e-4) >= 0.95
    c.close()


def test_precision_argument_is_validated(built_lib):
    with pytest.raises(abi.LevinsError):
        abi.Context(0, deskew_precision=2)
