"""Batched plane geometry of lidar::PointCloudCommon and the depth-factor / chi-square batches of the
optimiser (rank 4) on the GPU, against the oracle's restatements (pointcloud.cc:73-190, depth_factor.h:38-42,
optimizer.cc:659-665, 752-783), which tests/test_oracle.py and tests/test_host_mirror.py pin against numpy."""
import numpy as np
import pytest

from levins import synth

pytestmark = pytest.mark.gpu


def _patches(n_groups, seed):
    """groups of 5 nearly coplanar points (two thirds planar to 1 cm, the rest scattered 0.4 m off the plane)"""
    rng = np.random.default_rng(seed)
    pts = np.zeros(5 * n_groups, synth.XYZI)
    base = rng.normal(0, 8, (n_groups, 3)) + [0, 0, 15]
    nrm = rng.normal(0, 1, (n_groups, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    for g in range(n_groups):
        off = rng.normal(0, 0.4, (5, 3))
        if g % 3:
            off -= np.outer(off @ nrm[g], nrm[g]) * 0.975
        p = (base[g] + off).astype(np.float32)
        pts["x"][5 * g:5 * g + 5], pts["y"][5 * g:5 * g + 5], pts["z"][5 * g:5 * g + 5] = p[:, 0], p[:, 1], p[:, 2]
    pts["w"] = 1.0
    return pts


def test_plane_estimation_batch(ctx, oracle):
    pts = _patches(3000, 1)
    got = ctx.plane_estimation_batch(pts, 0.1)
    n_ok = 0
    for g in range(0, 3000, 7):
        ok, n, ninv, dist = oracle.plane_estimation(pts[5 * g:5 * g + 5], 0.1)
        assert bool(got["valid"][g]) == ok and np.array_equal(got["unit_normal"][g], n) and got["norm_inverse"][g] == ninv
        assert got["n_distance"][g] == len(dist) and np.array_equal(got["distance"][g, :len(dist)], dist)
        n_ok += ok
    assert 50 < n_ok < 400  # both outcomes
    assert 0.5 < got["valid"].mean() < 0.8


def test_point_to_plane_std_batch(ctx, oracle):
    pts = _patches(2000, 2)
    rng = np.random.default_rng(3)
    targets = pts[rng.integers(0, len(pts), 2000)].copy()
    got = ctx.point_to_plane_std_batch(pts, 5, targets, 0.03)
    for g in range(0, 2000, 9):
        want = oracle.point_to_plane_std(pts[5 * g:5 * g + 5], targets[g], 0.03)
        assert np.isclose(got[g], want, rtol=1e-12, atol=0), (g, got[g], want)
    assert np.all(np.isfinite(got)) and np.all(got > 0)


def test_depth_factor_and_chi2_batches(ctx, oracle):
    rng = np.random.default_rng(4)
    n = 5000
    depth = rng.uniform(1.0, 200.0, n)
    invdepth = 1.0 / (depth + rng.normal(0, 0.3, n) * rng.choice([0.1, 1.0, 10.0], n))
    got = ctx.depth_factor_batch(depth, invdepth, sigma=0.1, huber_a=1.0)
    for k in range(0, n, 11):
        r, j, c = oracle.depth_factor(depth[k], invdepth[k], 0.1, 1.0)
        assert got["residual"][k] == r and got["jacobian"][k] == j and got["cost"][k] == c
    # independent restatement: the residual's derivative by central differences, Huber cost by its definition
    h = 1e-7 * invdepth
    num = ((depth - 1.0 / (invdepth + h)) - (depth - 1.0 / (invdepth - h))) / 0.1 / (2 * h)
    assert np.allclose(got["jacobian"], num, rtol=1e-5)
    s = got["residual"] ** 2
    assert np.allclose(got["cost"], 0.5 * np.where(s <= 1.0, s, 2 * np.sqrt(s) - 1.0), rtol=1e-14)
    assert 0.05 < np.mean(s > 1.0) < 0.95  # both branches of the loss
    plain = ctx.depth_factor_batch(depth, invdepth, sigma=0.1, huber_a=0.0)
    assert np.array_equal(plain["cost"], 0.5 * plain["residual"] ** 2)
    res = rng.normal(0, 1.3, (n, 2))
    cost, outlier = ctx.chi2_culling_batch(res, 5.991)
    for k in range(0, n, 13):
        o, c = oracle.chi2_outlier(res[k], 5.991)
        assert bool(outlier[k]) == o and cost[k] == c
    assert 0.02 < outlier.mean() < 0.5
