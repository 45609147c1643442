"""Seeded differential fuzzing of the fused keyframe path against the oracle: random rigs, motions, INS rates (regular and
jittered), time offsets, window lengths, point distributions (inside / outside the FOV, around the range gates, clustered
on a few cells) and feature sets.  Every scenario checks window membership, range-gate decisions, coordinates, M and the
per-feature decisions."""
import os

import numpy as np
import pytest

from levins import abi, synth

pytestmark = pytest.mark.gpu


def _rand_rot(rng, max_angle):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    a = rng.uniform(-max_angle, max_angle)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def _scenario(seed):
    rng = np.random.default_rng(seed)
    t0 = float(rng.uniform(1e3, 6e5))
    interval = float(rng.choice([0.2, 0.5, 1.0]))
    stamp = t0 + 2.0 + rng.uniform(0, 0.01)
    n = int(rng.integers(200, 40_000)) * (4 if seed % 2 == 1 else 1)
    # points: time-sorted over a span a little longer than the window on both sides
    times = np.sort(rng.uniform(stamp - interval - 0.2, stamp + 0.1, n))
    kind = rng.integers(0, 3)
    if kind == 0:      # all directions
        d = rng.normal(size=(n, 3))
    elif kind == 1:    # mostly in front of the camera-ish axis
        d = rng.normal(size=(n, 3)) * [0.5, 0.5, 0.2] + [0, 0, 1]
    else:              # clustered: many points per cell
        d = rng.normal(size=(n, 3)) * 0.03 + rng.normal(size=(1, 3)) * [0.3, 0.3, 0] + [0, 0, 1]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    planar = seed % 2 == 1  # every other scenario: a wall (slow motion below), so that plane fits succeed and stage 3 decides
    if planar:
        nrm = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), 1.0])
        nrm /= np.linalg.norm(nrm)
        dist = rng.uniform(3.0, 40.0)
        cosang = d @ nrm
        r = np.where(cosang > 0.05, dist / np.maximum(cosang, 0.05), rng.uniform(2.0, 100.0, n))
        r = r + rng.normal(0, rng.choice([0.0, 0.01, 0.05]), n)          # sensor noise: some fits exceed the 0.1 m residual
        out = rng.random(n) < 0.05
        r[out] = r[out] * rng.uniform(0.5, 1.5, out.sum())                # depth discontinuities: range-spread rejections
        r = np.clip(r, 0.3, 400.0)
    else:
        r = np.exp(rng.uniform(np.log(0.5), np.log(260.0), n))  # both range gates are exercised
    r[rng.random(n) < 0.02] = 1.0
    r[rng.random(n) < 0.02] = 200.0
    pts = np.zeros(n, synth.XYZIT)
    R_lc = _rand_rot(rng, 0.3)  # points are generated roughly in the camera frame, then moved into the lidar frame
    p = (d * r[:, None]) @ R_lc.T
    pts["x"], pts["y"], pts["z"] = p[:, 0].astype(np.float32), p[:, 1].astype(np.float32), p[:, 2].astype(np.float32)
    pts["intensity"] = rng.uniform(0, 255, n).astype(np.float32)
    pts["time"] = times
    # extrinsics
    pose_c_l = synth.pose_to_abi(R_lc.T @ _rand_rot(rng, 0.05), rng.normal(size=3) * 0.2)
    pose_b_c = synth.pose_to_abi(_rand_rot(rng, np.pi), rng.normal(size=3) * 0.5)
    td = float(rng.choice([0.0, -0.48, 0.013]))
    # INS: random rate, optional jitter and gaps, window possibly shorter than the stack on either side
    rate = float(rng.choice([20.0, 100.0, 200.0, 1000.0]))
    a, b = stamp - interval - rng.uniform(-0.1, 0.4) + td, stamp + rng.uniform(-0.05, 0.3) + td
    tt = np.arange(a, b, 1.0 / rate)
    if rng.random() < 0.5:
        tt = np.sort(tt + rng.uniform(-0.3, 0.3, len(tt)) / rate)
    if rng.random() < 0.3 and len(tt) > 20:
        tt = np.delete(tt, rng.choice(np.arange(1, len(tt) - 1), len(tt) // 4, replace=False))
    tt = tt[:4000]
    ins = np.zeros(len(tt), synth.INS)
    ins["time"] = tt
    ins["state_time"] = tt + rng.uniform(-1e-4, 1e-4)
    w = rng.normal(size=3) * (0.05 if planar else rng.choice([0.1, 1.0, 5.0]))  # angular rate, rad/s
    v = rng.normal(size=3) * (0.2 if planar else rng.choice([0.5, 5.0, 20.0]))
    ang = np.linalg.norm(w) * (tt - tt[0])
    ax = w / np.linalg.norm(w)
    ins["q"][:, :3] = np.sin(0.5 * ang)[:, None] * ax
    ins["q"][:, 3] = np.cos(0.5 * ang)
    ins["p"] = v[None, :] * (tt - tt[0])[:, None] + 0.05 * np.sin(3.0 * (tt - tt[0]))[:, None]
    # features: unit rays inside the FOV
    f = int(rng.integers(1, 200))
    q = np.stack([rng.uniform(-0.8, 0.8, f), rng.uniform(-0.8, 0.8, f), np.ones(f)], axis=1)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    return dict(pts=pts, ins=ins, stamp=stamp, interval=interval, pose_b_c=pose_b_c, pose_c_l=pose_c_l, td=td, q=q)


N_SEEDS = int(os.environ.get("LEVINS_FUZZ_SEEDS", "24"))  # e.g. LEVINS_FUZZ_SEEDS=400 for a soak run


@pytest.mark.parametrize("seed", range(N_SEEDS))
def test_fuzz_keyframe_against_oracle(oracle, built_lib, seed):
    s = _scenario(1000 + seed)
    c = abi.Context(0, stack_interval=s["interval"])
    c.set_extrinsics(s["pose_b_c"], s["pose_c_l"], s["td"])
    c.append_scan(s["pts"])
    ref = oracle.stack_deskew(s["pts"], s["stamp"], s["ins"], s["pose_b_c"], s["pose_c_l"], s["td"], s["interval"], details=True)
    st, n, cloud = c.stack_deskew(s["stamp"], s["ins"])
    assert n == ref["n"], (n, ref["n"])
    if n:
        kept = ref["kept"].astype(bool)
        a = np.stack([cloud["x"], cloud["y"], cloud["z"]], axis=-1)
        b = np.stack([ref["cloud"]["x"], ref["cloud"]["y"], ref["cloud"]["z"]], axis=-1)
        assert np.array_equal(np.any(a != 0, axis=1) | ~kept, np.ones(n, bool)) or np.array_equal(np.any(a != 0, axis=1), kept)
        if kept.any():
            err = np.linalg.norm(a[kept].astype(np.float64) - b[kept], axis=1) / np.maximum(np.linalg.norm(b[kept].astype(np.float64), axis=1), 1e-30)
            assert err.max() <= 1e-4, err.max()
            assert np.mean(np.all(a[kept] == b[kept], axis=1)) >= 0.999
    c.buffer_clear()
    c.append_scan(s["pts"])
    got = c.keyframe(s["stamp"], s["ins"], s["q"])
    want = oracle.keyframe(s["pts"], s["stamp"], s["ins"], s["pose_b_c"], s["pose_c_l"], s["td"], s["q"], s["interval"])
    assert got["n_projected"] == want["n_projected"] and got["m"] == want["m"], (got["m"], want["m"])
    assert np.array_equal(got["type"], want["type"])
    both = got["type"] == 1
    if both.any():
        d0, d1 = want["xyz_depth"][both, 3].astype(np.float64), got["xyz_depth"][both, 3].astype(np.float64)
        assert np.max(np.abs(d1 - d0) / np.abs(d0)) <= 1e-4
    c.close()
