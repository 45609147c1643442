"""The single-shot pose lookups the library exports for the host (levins_ins_window_index, levins_state_interpolate,
levins_state_to_pose, levins_pose_at; MISC of the reference, common/misc.cc:25-103).  They run on the CPU (same source as the
device code in k_prepare), so they are checked here without a GPU: against the oracle bit for bit and against scipy."""
import ctypes as C

import numpy as np
import pytest

from levins import abi, synth


def _window(rng, n, kind):
    t = 345600.0 + np.arange(n) / 200.0
    if kind == "jitter":
        t = np.sort(t + rng.uniform(-0.3, 0.3, n) / 200.0)
    elif kind == "gaps":
        t = np.delete(t, rng.choice(np.arange(1, n - 1), n // 4, replace=False))
    ins = np.zeros(len(t), synth.INS)
    ins["time"] = t
    ins["state_time"] = t + (rng.uniform(-1e-4, 1e-4) if kind != "uniform" else 0.0)
    w = rng.normal(size=3) * 0.8
    ang = np.linalg.norm(w) * (t - t[0])
    ins["q"][:, :3] = np.sin(0.5 * ang)[:, None] * (w / np.linalg.norm(w))
    ins["q"][:, 3] = np.cos(0.5 * ang)
    ins["p"] = rng.normal(size=3)[None, :] * (t - t[0])[:, None] + 0.1 * np.sin(2.0 * (t - t[0]))[:, None]
    return ins


@pytest.mark.parametrize("kind", ["uniform", "jitter", "gaps"])
def test_pose_lookup_exports_equal_the_oracle(oracle, built_lib, kind):
    L = abi.lib()
    L.levins_ins_window_index.restype = C.c_size_t
    L.levins_ins_window_index.argtypes = [C.c_void_p, C.c_size_t, C.c_double]
    L.levins_pose_at.restype = C.c_int
    L.levins_pose_at.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]
    L.levins_state_interpolate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    L.levins_state_to_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng({"uniform": 1, "jitter": 2, "gaps": 3}[kind])
    ins = _window(rng, 700, kind)
    ext = synth.rig_r3live().pose_b_c
    t = ins["time"]
    times = np.concatenate([rng.uniform(t[0] - 0.01, t[-1] + 0.01, 400), t[[0, 1, 50, -2, -1]], [t[0] - 5.0, t[-1] + 5.0]])
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for tm in times:
        idx = L.levins_ins_window_index(p(ins), len(ins), float(tm))
        assert idx == oracle.get_ins_window_index(ins, tm)
        pose = np.zeros(12)
        ok = L.levins_pose_at(p(ins), len(ins), p(ext), float(tm), p(pose))
        want_ok, want = oracle.get_pose_from_ins_window(ins, ext, tm)
        assert bool(ok) == want_ok and np.array_equal(pose, want)
        if idx > 0:
            out = np.zeros(1, synth.INS)
            L.levins_state_interpolate(p(ins[idx - 1:idx]), p(ins[idx:idx + 1]), float(tm), p(out))
            pp, qq = oracle.state_pose_interpolation(ins[idx - 1], ins[idx], tm)
            assert np.array_equal(out["p"][0], pp) and np.array_equal(out["q"][0], qq) and out["time"][0] == tm
            pose2 = np.zeros(12)
            L.levins_state_to_pose(p(out), p(ext), p(pose2))
            assert np.array_equal(pose2, pose)


def test_pose_lookup_against_scipy(built_lib):
    """independent check: SLERP + linear position, then the extrinsic"""
    from scipy.spatial.transform import Rotation, Slerp

    L = abi.lib()
    L.levins_pose_at.restype = C.c_int
    L.levins_pose_at.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]
    rng = np.random.default_rng(7)
    ins = _window(rng, 300, "uniform")
    rig = synth.rig_r3live()
    ext = rig.pose_b_c
    sl = Slerp(ins["time"], Rotation.from_quat(ins["q"]))
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for tm in rng.uniform(ins["time"][1], ins["time"][-2], 100):
        pose = np.zeros(12)
        assert L.levins_pose_at(p(ins), len(ins), p(ext), float(tm), p(pose)) == 1
        Rq = sl([tm]).as_matrix()[0]
        pos = np.array([np.interp(tm, ins["time"], ins["p"][:, i]) for i in range(3)])
        assert np.allclose(pose[:9].reshape(3, 3).T, Rq @ rig.R_b_c, atol=1e-12)
        assert np.allclose(pose[9:], pos + Rq @ rig.t_b_c, atol=1e-9)
