"""CPU checks of the "exact by construction" arguments of the kernels (csrc/device_math.cuh), restated in numpy with f32
arithmetic.  The GPU tests check the kernels' results; these check the arguments themselves.

1. outside_fov: the FOV gate of tracking.cc:311-314, fl32(|x| / |z|) > tan(45 deg) (= 0x1.fffffffffffffp-1 as a double),
   restated without the division as |x| - |z| >= -|z| 2^-25.
2. the polynomial cell path: its cell coordinate is within cell_guard / 2 of the exact one, so every id it decides without the
   exact path (distance to a rounding boundary > cell_guard) is the reference's.
3. the one-compare bracket test of the interval entries.
4. the 9 x 13 window and the cell mask of k_associate hold every point within the acceptance radius of a query.
5. the range gate restated on the squared norm (two f32 thresholds found by bisection).
6. the 64-bit z-buffer key: its minimum is the reference's winner in any arrival order; stale generations read as empty.
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


# ---- 4. the windowed 5-NN of k_associate: the 9 x 13 window and the cell mask never drop a point that can matter ----
def _query_cell_position(q):
    """numpy twin of k_associate's query position in cell units (kernels.cuh, "The query's position in cell units")."""
    qx, qy, qz = q[:, 0], q[:, 1], q[:, 2]
    inv_ang = F(1.0 / ANGLE)
    off = F(np.pi * 0.5 - 90.0 * np.pi / 180.0 * 0.5)
    inside = (qz > 0) & (np.abs(qx) <= qz) & (np.abs(qy) <= qz)
    with np.errstate(divide="ignore", invalid="ignore"):
        el_in = _atan_poly((qy / np.sqrt(_fma(qx, qx, (qz * qz).astype(F))).astype(F)).astype(F))
        az_in = _atan_poly((qx / qz).astype(F))
    rq_in = _fma(el_in, np.full_like(el_in, inv_ang), np.full_like(el_in, F(150.0)))
    cq_in = _fma(-az_in, np.full_like(el_in, inv_ang), np.full_like(el_in, F(150.0)))
    el_out = np.arctan2(qy, np.sqrt((qx * qx + qz * qz).astype(F)).astype(F)).astype(F)
    az_out = np.arctan2(qz, qx).astype(F)
    rq_out = ((el_out + off).astype(F) * inv_ang).astype(F)
    cq_out = ((az_out - off).astype(F) * inv_ang).astype(F)
    rq = np.where(inside, rq_in, rq_out).astype(F)
    cq = np.where(inside, cq_in, cq_out).astype(F)
    ael = np.abs(np.where(inside, el_in, el_out)).astype(F)
    return rq, cq, ael


def _reference_cell(p):
    """tracking.cc:311-328 on f32 points: (kept, row, col)"""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        gate = (z < 0) | (np.abs((x / z).astype(F)).astype(np.float64) > TAN) | (np.abs((y / z).astype(F)).astype(np.float64) > TAN)
    half = np.pi * 0.5 - 90.0 * np.pi / 180.0 * 0.5
    s = (x * x + z * z).astype(F)
    row_angle = np.arctan2(y, np.sqrt(s).astype(F)).astype(F).astype(np.float64) + half
    col_angle = np.arctan2(z, x).astype(F).astype(np.float64) - half
    row = np.round(row_angle / ANGLE).astype(np.int64)
    col = np.round(col_angle / ANGLE).astype(np.int64)
    kept = ~gate & (row >= 0) & (col >= 0) & (row < 300) & (col < 300)
    return kept, row, col


def _rotate_about(q, axis, g):
    """Rodrigues: rotate the unit vectors q by the angles g about the unit axes `axis` (f64)"""
    c, s = np.cos(g)[:, None], np.sin(g)[:, None]
    return q * c + np.cross(axis, q) * s + axis * np.sum(axis * q, axis=1, keepdims=True) * (1 - c)


def test_association_window_and_mask_keep_every_point_within_the_acceptance_radius():
    """tracking.cc:573 accepts a feature only if its 5th neighbour lies within the chord T = 3 sin(0.3 deg), so only points
    within T of the query can influence a result.  Every such point that the range image can hold must lie (a) inside the
    9 x 13 window around the query's cell and (b) inside the mask  dr^2 + cmin dc^2 <= 3.01^2  that k_associate applies to
    the window's cells.  Queries inside and outside the 90-degree cone, points up to exactly T away, emphasis on the
    corners of the image where the azimuth stretch 1 / cos(el) is largest."""
    rng = np.random.default_rng(11)
    n = 3_000_000
    # query rays: u = x / z, v = y / z over a range a little wider than the image (camera FOV can exceed the range image)
    u = rng.uniform(-1.25, 1.25, n)
    v = rng.uniform(-1.25, 1.25, n)
    edge = rng.random(n) < 0.4  # concentrate on the top / bottom rim, where cos(el) is smallest
    v = np.where(edge, np.sign(v) * rng.uniform(0.9, 1.05, n), v)
    u = np.where(edge & (rng.random(n) < 0.5), rng.uniform(-0.2, 0.2, n), u)
    q64 = np.stack([u, v, np.ones(n)], axis=1)
    q64 /= np.linalg.norm(q64, axis=1, keepdims=True)
    q = q64.astype(F)  # the query the kernel sees (f32 unit ray, tracking.cc:395)
    qn = q.astype(np.float64)
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    # points: rotate the query by an arc g with chord 2 sin(g / 2) <= T about a random axis perpendicular to it
    T = 3.0 * np.sin(ANGLE)
    gmax = 2.0 * np.arcsin(T / 2.0)
    g = gmax * np.where(rng.random(n) < 0.5, 1.0 - 1e-9, np.sqrt(rng.random(n)))
    t = rng.normal(size=(n, 3))
    t -= qn * np.sum(t * qn, axis=1, keepdims=True)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    axis = np.cross(qn, t)
    p_unit = _rotate_about(qn, axis, g)
    rho = np.exp(rng.uniform(np.log(1.0), np.log(150.0), n))
    p = (p_unit * rho[:, None]).astype(F)  # the deskewed point the range image bins (f32 camera-frame coordinates)
    kept, row, col = _reference_cell(p)
    # the unit-sphere point the search measures (tracking.cc:348-349) against the query, in f32 like L2_Simple<float>
    pu = (p / np.sqrt(np.sum(p * p, axis=1, dtype=F)).astype(F)[:, None]).astype(F)
    d = (pu - q).astype(F)
    d2 = np.sum(d * d, axis=1, dtype=F)
    matters = kept & (d2.astype(np.float64) <= T * T)
    assert matters.sum() > 1_000_000

    for poly_err in (0.0, +0.004, -0.004):  # the polynomial's error budget in cell units (see the test above)
        rq, cq, ael = _query_cell_position(q)
        rq = (rq + F(poly_err)).astype(F)
        cq = (cq - F(poly_err)).astype(F)
        rowq, colq = np.rint(rq).astype(np.int64), np.rint(cq).astype(np.int64)
        drow, dcol = np.abs(row - rowq)[matters], np.abs(col - colq)[matters]
        assert drow.max() <= 4 and dcol.max() <= 6, (int(drow.max()), int(dcol.max()))
        cmin = (np.maximum(F(0), np.cos(ael) * np.cos(np.minimum(ael + F(0.016), F(1.5707963)))) * F(0.999)).astype(F)
        dr = np.maximum(F(0), np.abs(row.astype(F) - rq) - F(0.51)).astype(F)
        dc = np.maximum(F(0), np.abs(col.astype(F) - cq) - F(0.51)).astype(F)
        inmask = (dr * dr + cmin * dc * dc).astype(F) <= F(9.0601)
        assert inmask[matters].all(), int((~inmask[matters]).sum())
    # the window is not oversized by much: offsets of 3-4 rows and 5 columns do occur, and the mask's left side reaches
    # 8.96 of the 9.06 it allows
    assert drow.max() >= 3 and dcol.max() >= 5


# ---- 5. the range gate on the squared norm ----
def _bits_to_f32(b):
    return np.array(b, dtype=np.uint32).view(F)


def _gate_thresholds(nearest, farthest):
    """numpy twin of make_consts (csrc/levins_abi.cu): bisection over the ordered positive f32 patterns for the largest n2 that
    is not "too far" and the smallest n2 that is not "too near" under  (double) sqrtf(n2) > farthest / < nearest."""
    def root(b):
        return float(np.sqrt(_bits_to_f32(b)))  # numpy's f32 sqrt is correctly rounded, like sqrtf and the device's sqrt.rn

    lo, hi = 0, 0x7F7FFFFF
    if root(hi) > farthest:
        while hi - lo > 1:
            mid = lo + (hi - lo) // 2
            if root(mid) > farthest:
                hi = mid
            else:
                lo = mid
        n2_far = F(-1.0) if root(lo) > farthest else _bits_to_f32(lo)
    else:
        n2_far = _bits_to_f32(hi)
    lo, hi = 0, 0x7F7FFFFF
    if root(lo) < nearest:
        while hi - lo > 1:
            mid = lo + (hi - lo) // 2
            if root(mid) < nearest:
                lo = mid
            else:
                hi = mid
        n2_near = _bits_to_f32(hi)
    else:
        n2_near = F(0.0)
    return F(n2_near), F(n2_far)


def test_range_gate_on_the_squared_norm_is_the_reference_gate():
    """le_vins.cc:447-451: norm = getVector3fMap().norm() (f32 sqrt of the f32 squared norm), gated when
    norm > FARTHEST_DEPTH or norm < NEAREST_DEPTH with the compare in double.  The kernels compare the squared norm with two
    f32 thresholds instead; both forms agree on every f32 within +-2^16 patterns of each threshold and on a broad sample,
    for the reference's constants and for thresholds that are not squares of floats."""
    rng = np.random.default_rng(5)
    for nearest, farthest in ((1.0, 200.0), (0.3, 120.5), (1.0 / 3.0, 77.7), (0.0, 1e30)):
        n2_near, n2_far = _gate_thresholds(nearest, farthest)
        around = []
        for thr in (n2_near, n2_far):
            b = int(np.array(thr, dtype=F).view(np.uint32))
            around.append(np.arange(max(b - 65536, 0), min(b + 65536, 0x7F7FFFFF), dtype=np.int64).astype(np.uint32).view(F))
        broad = np.exp(rng.uniform(np.log(1e-8), np.log(1e9), 2_000_000)).astype(F)
        n2 = np.concatenate(around + [broad, np.zeros(1, F)])
        norm = np.sqrt(n2).astype(np.float64)
        reference = (norm > farthest) | (norm < nearest)
        restated = (n2 > n2_far) | (n2 < n2_near)
        assert np.array_equal(reference, restated), (nearest, farthest, int(np.sum(reference != restated)))
        if farthest < 1e20:
            assert 0.05 < reference.mean() < 0.95


# ---- 6. the z-buffer key ----
KEY_IDX_BITS = 25


def _key_make(gen, rng_f32, idx):
    return (np.uint64(gen) << np.uint64(56)) | (rng_f32.view(np.uint32).astype(np.uint64) << np.uint64(KEY_IDX_BITS)) | idx.astype(np.uint64)


def test_zbuffer_key_minimum_is_the_reference_winner_and_stale_generations_read_as_empty():
    """device_math.cuh key_make: [gen:8][f32 range bits:31][point index:25].  (a) The minimum key of a cell is the point the
    reference's sequential loop keeps (tracking.cc:332: a strictly smaller f32 range replaces, so on equal ranges the first
    index stays), whatever the order in which the points arrive.  (b) gen counts down from 254 (next_gen, levins_abi.cu), so
    any key of the current use is smaller than every key a previous use left behind, and a cell whose minimum still carries
    an older gen (or the 0xFF.. clear pattern) is empty."""
    rng = np.random.default_rng(9)
    n, cells = 400_000, 5_000
    cell = rng.integers(0, cells, n)
    # few distinct ranges per cell so that ties are common; include denormal-free tiny and large values
    rngs = (rng.integers(1, 40, n).astype(F) * F(0.37)).astype(F)
    rngs[rng.random(n) < 0.01] = F(1e-30)
    rngs[rng.random(n) < 0.01] = F(3e38)
    idx = np.arange(n, dtype=np.int64)
    assert n < (1 << KEY_IDX_BITS) and np.all(rngs.view(np.uint32) < (1 << 31))

    # reference: sequential, strictly-smaller replaces
    best = np.full(cells, np.finfo(F).max, F)
    win = np.full(cells, -1, np.int64)
    for k in range(n):
        c = cell[k]
        if rngs[k] < best[c]:
            best[c] = rngs[k]
            win[c] = k

    gen = 200
    stale = _key_make(gen + 1, (rng.random(cells).astype(F) * F(1e-3)).astype(F), rng.integers(0, 1 << KEY_IDX_BITS, cells))  # nearer and older
    stale[::7] = np.uint64(0xFFFFFFFFFFFFFFFF)  # cells nobody touched since the clear
    for order in (np.arange(n), rng.permutation(n), np.arange(n)[::-1]):
        z = stale.copy()
        np.minimum.at(z, cell[order], _key_make(gen, rngs[order], idx[order]))  # what atomicMin leaves, in any arrival order
        live = (z >> np.uint64(56)) == np.uint64(gen)
        assert np.array_equal(live, win >= 0)
        got = (z & np.uint64((1 << KEY_IDX_BITS) - 1)).astype(np.int64)
        assert np.array_equal(got[live], win[live])
        got_range = ((z >> np.uint64(KEY_IDX_BITS)) & np.uint64(0x7FFFFFFF)).astype(np.uint32).view(F)
        assert np.array_equal(got_range[live], best[live])
    # an untouched cell of this use: the stale key stays and is not live
    z = stale.copy()
    assert not np.any((z >> np.uint64(56)) == np.uint64(gen))
    # the generation sequence: 254 after a clear, then down to 1, then clear again -- never 0xFF (the clear pattern) or 0
    g, seen = 0, []
    for _ in range(600):
        g = 254 if (g <= 1 or g > 254) else g - 1
        seen.append(g)
    assert min(seen) == 1 and max(seen) == 254 and seen[0] == 254 and seen[253] == 1 and seen[254] == 254


# ---- the twins above restate THESE sources: a changed constant in the kernels must fail here, not pass silently ----
def test_numpy_twins_use_the_constants_of_the_cuda_sources():
    import os
    import re

    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "le-vins_b200", "csrc")
    dm = open(os.path.join(csrc, "device_math.cuh")).read()
    kn = open(os.path.join(csrc, "kernels.cuh")).read()
    ab = open(os.path.join(csrc, "levins_abi.cu")).read()

    body = dm[dm.index("float atan_poly(float t)"):]
    body = body[:body.index("return t * p;")]
    coeffs = [float(c) for c in re.findall(r"(-?\d+\.\d+(?:e-?\d+)?)f", body)]
    twin = [0.0024566929787397385, -0.014401225373148918, 0.03978099301457405, -0.07234836369752884, 0.10498934984207153,
            -0.14161226153373718, 0.19985906779766083, -0.33332598209381104, 0.9999998807907104]
    assert coeffs == twin

    assert re.search(r"constexpr int WIN_ROWS = 4;", kn) and re.search(r"constexpr int WIN_COLS = 6;", kn)
    for lit in ("- 0.51f)", "<= 9.0601f", "ael + 0.016f", "* 0.999f", "150.0f"):
        assert lit in kn, lit
    assert "c.cell_guard = 4e-4f;" in ab and CELL_GUARD == 4e-4
    assert "constexpr unsigned KEY_IDX_BITS = 25;" in dm and KEY_IDX_BITS == 25
    assert "(unsigned long long) gen << 56" in dm
    assert "*gen = 254;" in ab and "if (*gen <= 1 || *gen > 254)" in ab
    assert "2.98023223876953125e-8f" in dm  # 2^-25 of the division-free FOV gate
