"""CPU checks of two "exact by construction" arguments of the kernels (csrc/device_math.cuh), restated in numpy with f32
arithmetic.  The GPU tests check the kernels' results; these check the arguments themselves.

1. outside_fov: the FOV gate of tracking.cc:311-314, fl32(|x| / |z|) > tan(45 deg) (= 0x1.fffffffffffffp-1 as a double),
   restated without the division as |x| - |z| >= -|z| 2^-25.
2. the polynomial cell path: its cell coordinate is within cell_guard / 2 of the exact one, so every id it decides without the
   exact path (distance to a rounding boundary > cell_guard) is the reference's.
"""
import numpy as np

F = np.float32
TAN = np.tan(90.0 * np.pi / 180.0 * 0.5)  # tracking.h:128, evaluates to 0x1.fffffffffffffp-1
ANGLE = 0.3 * np.pi / 180.0
CELL_GUARD = 4e-4


def test_tan_constant():
    assert TAN == float.fromhex("0x1.fffffffffffffp-1")


def test_fov_gate_without_division_is_exact():
    rng = np.random.default_rng(0)
    n = 2_000_000
    z = np.exp(rng.uniform(np.log(1e-3), np.log(300.0), n)).astype(F)
    # |x| within a few ulps of |z| (where the two formulations could differ), plus a broad background
    k = rng.integers(-6, 7, n)
    x = z.copy()
    for step in range(6):
        up = k > step
        dn = k < -step
        x = np.where(up, np.nextafter(x, F(np.inf)), np.where(dn, np.nextafter(x, F(0)), x)).astype(F)
    bg = rng.random(n) < 0.3
    x = np.where(bg, (z * rng.uniform(0.0, 2.0, n)).astype(F), x).astype(F)
    x = np.where(rng.random(n) < 0.5, -x, x).astype(F)
    reference = np.abs((x / z).astype(F)).astype(np.float64) > TAN  # fabs(x / z) > TAN with the division in f32
    az = np.abs(z)
    lim = (-az * F(2.98023223876953125e-8)).astype(F)
    restated = (np.abs(x) - az).astype(F) >= lim
    assert np.array_equal(reference, restated), int(np.sum(reference != restated))
    assert 0.2 < reference.mean() < 0.8


def _fma(a, b, c):
    """f32 fused multiply-add: the product of two f32 is exact in f64, one rounding at the end"""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def _atan_poly(t):
    u = (t * t).astype(F)
    p = np.full_like(t, F(0.0024566929787397385))
    for c in (-0.014401225373148918, 0.03978099301457405, -0.07234836369752884, 0.10498934984207153, -0.14161226153373718,
              0.19985906779766083, -0.33332598209381104, 0.9999998807907104):
        p = _fma(p, u, np.full_like(t, F(c)))
    return (t * p).astype(F)


def test_polynomial_cell_coordinate_is_within_half_the_guard():
    rng = np.random.default_rng(1)
    n = 1_000_000
    # camera-frame points inside the 90-degree cone, ranges 1 .. 200 m
    d = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), np.ones(n)], -1)
    p = (d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(1, 200, (n, 1))).astype(F)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    worst = 0.0
    for sign in (-1.0, 1.0):  # MUFU.RCP / MUFU.RSQ are approximations: take both extremes of their 2^-22 relative error
        rz = ((1.0 / z.astype(np.float64)) * (1.0 + sign * 2.0 ** -22)).astype(F)
        s = _fma(x, x, (z * z).astype(F))
        rs = ((1.0 / np.sqrt(s.astype(np.float64))) * (1.0 + sign * 2.0 ** -22)).astype(F)
        col = _fma(_atan_poly((x * rz).astype(F)), np.full_like(x, F(-1.0 / ANGLE)), np.full_like(x, F(150.0)))
        row = _fma(_atan_poly((y * rs).astype(F)), np.full_like(x, F(1.0 / ANGLE)), np.full_like(x, F(150.0)))
        # the exact cell coordinates (what tracking.cc:317-324 rounds), in f64 from the same f32 inputs
        xd, yd, zd = x.astype(np.float64), y.astype(np.float64), z.astype(np.float64)
        col_exact = (np.arctan2(zd, xd) - np.pi / 4) / ANGLE
        row_exact = (np.arctan2(yd, np.sqrt(xd * xd + zd * zd)) + np.pi / 4) / ANGLE
        worst = max(worst, float(np.max(np.abs(col - col_exact))), float(np.max(np.abs(row - row_exact))))
    # the reference itself evaluates the angles in f32 (atan2f): up to ~1e-7 rad = 2e-5 cell away from the f64 value
    assert worst + 2e-5 < CELL_GUARD / 2, worst


def test_one_compare_bracket_test():
    """key_brackets (device_math.cuh): for positive doubles t_lo <= t < t_hi is ONE unsigned 64-bit compare on the bit patterns,
    (bits(t) - bits(t_lo)) < (bits(t_hi) - bits(t_lo)); with the fallback entry's (t_lo, t_hi) = (last sample, first sample) the
    width wraps and the same compare is true exactly for the times OUTSIDE [first, last), negative times included."""
    rng = np.random.default_rng(2)
    n = 1_000_000
    lo = rng.uniform(1e3, 6e5, n)
    hi = lo + np.exp(rng.uniform(np.log(1e-4), np.log(10.0), n))
    t = np.where(rng.random(n) < 0.5, rng.uniform(lo - 1.0, hi + 1.0), np.where(rng.random(n) < 0.5, lo, hi))
    t = np.where(rng.random(n) < 0.02, -t, t)  # negative times wrap to huge unsigned values and miss
    t = np.where(rng.random(n) < 0.05, np.nextafter(hi, 0), t)
    b = lambda a: np.asarray(a, np.float64).view(np.uint64)  # noqa: E731
    with np.errstate(over="ignore"):
        inside = (b(t) - b(lo)) < (b(hi) - b(lo))
        assert np.array_equal(inside, (lo <= t) & (t < hi))
        first, last = lo, hi  # fallback entry: t_lo = last, t_hi = first
        outside = (b(t) - b(last)) < (b(first) - b(last))
        assert np.array_equal(outside, ~((first <= t) & (t < last)))
