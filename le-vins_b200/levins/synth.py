"""Seeded synthetic LiDAR / INS / feature generators for the depth-enhancement hot path.

Shapes follow SURVEY.md section 8(d) / BASELINE.json `configs`:

  cfg 1  Avia-like rosette, 24 000 pts/scan, 10 scans in the 0.5 s window, 752x480, F=150
  cfg 2  R3LIVE handheld: Avia 24 000 pts/scan @ 10 Hz (5 scans / window), 1280x1024, F=300
  cfg 3  FusionPortableV2 vehicle: Ouster OS1-128 (131 072 pts/scan), 20-scan window (2.0 s), 1440x1080, F=200
  cfg 4  VLP-16 sparse map (28 800 pts/scan, 5 scans), 1280x1024, F=500
  cfg 5  many independent keyframes of one of the shapes above (built with `make_batch`)

Everything is generated in the reference's conventions: body = forward/right/down, camera = z forward,
LiDAR = x forward / y left / z up, world z down; extrinsics are the values of the reference's own
config files (config/le_vins_r3live.yaml:60-61,71-72 and config/le_vins_fusionportable.yaml:62-63,70-71,75).
Points are produced by ray casting a static box world from the true (analytic) LiDAR pose at each point's
own timestamp, so that after a correct deskew the cloud lies on planes again.

Record layouts (numpy structured dtypes) are byte-identical to the C ABI structs in include/levins/abi.h.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

XYZIT = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("pad0", "<f4"), ("intensity", "<f4"), ("pad1", "<f4"), ("time", "<f8")]
)
XYZI = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"), ("intensity", "<f4"), ("pad", "<f4", (3,))])
INS = np.dtype([("time", "<f8"), ("state_time", "<f8"), ("p", "<f8", (3,)), ("q", "<f8", (4,))])
assert XYZIT.itemsize == 32 and XYZI.itemsize == 32 and INS.itemsize == 72

BASE_SEED = 20261019
T0 = 345600.0  # GPS seconds of week

# ----------------------------------------------------------------------------------------------
# small rotation helpers (no scipy dependency on the hot generation path)
# ----------------------------------------------------------------------------------------------


def quat_to_matrix(q):
    """q = (x, y, z, w) -> 3x3 (same convention as Eigen::Quaterniond::toRotationMatrix)."""
    x, y, z, w = [np.asarray(c, dtype=np.float64) for c in np.moveaxis(np.asarray(q, dtype=np.float64), -1, 0)]
    R = np.empty(x.shape + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def matrix_to_quat(R):
    """Batched rotation matrix -> quaternion (x, y, z, w), w >= 0."""
    from scipy.spatial.transform import Rotation

    q = Rotation.from_matrix(np.asarray(R)).as_quat()
    q = np.where(q[..., 3:4] < 0, -q, q)
    return q


def euler_zyx_matrix(yaw, pitch, roll):
    """R = Rz(yaw) Ry(pitch) Rx(roll), batched."""
    yaw, pitch, roll = np.broadcast_arrays(yaw, pitch, roll)
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    R = np.empty(yaw.shape + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def pose_to_abi(R, t):
    """(3x3, 3) -> 12 doubles: R column-major then t (Pose, le_vins/common/types.h:51-54)."""
    return np.concatenate([np.asarray(R, dtype=np.float64).T.reshape(-1), np.asarray(t, dtype=np.float64)]).copy()


# ----------------------------------------------------------------------------------------------
# sensor rig / trajectories
# ----------------------------------------------------------------------------------------------


@dataclass
class Rig:
    R_b_c: np.ndarray
    t_b_c: np.ndarray
    R_c_l: np.ndarray
    t_c_l: np.ndarray
    td_b_l: float
    intr: np.ndarray  # fx fy cx cy skew
    resolution: tuple

    @property
    def pose_b_c(self):
        return pose_to_abi(self.R_b_c, self.t_b_c)

    @property
    def pose_c_l(self):
        return pose_to_abi(self.R_c_l, self.t_c_l)


def rig_r3live(resolution=(1280, 1024), intr=(863.4241, 863.4171, 640.6808, 518.3392), td_b_l=0.0) -> Rig:
    return Rig(
        quat_to_matrix([0.499629, 0.504424, 0.495035, 0.500867]),
        np.array([0.058, -0.069, 0.013]),
        quat_to_matrix([-0.4991117338677414, 0.5062278362608855, -0.49440996678462257, -0.5001796073953708]),
        np.array([0.04183047861925253, 0.005597350405590019, -0.10409073314623786]),
        td_b_l,
        np.array(list(intr) + [0.0]),
        resolution,
    )


def rig_fusionportable(td_b_l=-0.48) -> Rig:
    s = 1440.0 / 1024.0  # BASELINE names 1440x1080; the repo's config is 1024x768 -> scale intrinsics
    return Rig(
        quat_to_matrix([0.494880, 0.498635, 0.504425, 0.502009]),
        np.array([0.385, -0.392, 0.048]),
        quat_to_matrix([-0.49990070625763156, 0.5025623801361645, -0.4972266202172702, -0.5002959385226958]),
        np.array([0.4041748623695919, -0.0921150811404132, -0.27708036513621204]),
        td_b_l,
        np.array([593.5669 * s, 593.8849 * s, 513.3945 * s, 401.9319 * s, 0.0]),
        (1440, 1080),
    )


@dataclass
class Trajectory:
    """Analytic body trajectory in a z-down world; body = forward / right / down."""

    speed: float = 1.0
    yaw_rate: float = 0.2
    wobble_deg: float = 2.0
    origin: tuple = (0.0, 0.0, 0.0)
    yaw0: float = 0.0

    def pose(self, t):
        t = np.asarray(t, dtype=np.float64) - T0
        yaw = self.yaw0 + self.yaw_rate * t
        w = np.deg2rad(self.wobble_deg)
        roll = w * np.sin(2 * np.pi * 0.7 * t)
        pitch = w * np.sin(2 * np.pi * 0.45 * t + 0.3)
        R = euler_zyx_matrix(yaw, pitch, roll)
        if abs(self.yaw_rate) > 1e-9:
            px = self.speed / self.yaw_rate * (np.sin(yaw) - np.sin(self.yaw0))
            py = -self.speed / self.yaw_rate * (np.cos(yaw) - np.cos(self.yaw0))
        else:
            px = self.speed * t * np.cos(self.yaw0)
            py = self.speed * t * np.sin(self.yaw0)
        pz = 0.05 * np.sin(2 * np.pi * 0.9 * t)
        p = np.stack([px + self.origin[0], py + self.origin[1], pz + self.origin[2]], axis=-1)
        return R, p


# ----------------------------------------------------------------------------------------------
# scene: axis-aligned boxes in the world; box 0 is the enclosing hall (hit from the inside)
# ----------------------------------------------------------------------------------------------


def make_scene(rng, kind="hall"):
    if kind == "hall":
        hall = np.array([[-40.0, -30.0, -12.0], [40.0, 30.0, 1.5]])
        n_box, spread, size = 14, (30.0, 22.0), (0.8, 4.0)
    else:  # street canyon
        hall = np.array([[-150.0, -14.0, -30.0], [150.0, 14.0, 1.9]])
        n_box, spread, size = 24, (140.0, 11.0), (1.0, 5.0)
    boxes = [hall]
    for _ in range(n_box):
        c = np.array([rng.uniform(-spread[0], spread[0]), rng.uniform(-spread[1], spread[1])])
        if np.hypot(c[0], c[1]) < 4.0:  # keep the start area free
            c += 6.0 * np.sign(c + 1e-3)
        sx, sy, sz = rng.uniform(size[0], size[1], size=3)
        if kind != "hall" and abs(c[1]) - sy / 2 < 3.0:  # street: keep the driving corridor |y| < 3 m free
            c[1] = np.sign(c[1] + 1e-3) * (3.0 + sy / 2 + rng.uniform(0.0, 6.0))
        boxes.append(np.array([[c[0] - sx / 2, c[1] - sy / 2, hall[1, 2] - sz], [c[0] + sx / 2, c[1] + sy / 2, hall[1, 2]]]))
    return np.stack(boxes)  # (B, 2, 3)


def raycast(boxes, o, d, chunk=1 << 18):
    """Nearest positive hit range of rays o + s d (N,3) against the boxes.  Slab method, float64."""
    n = o.shape[0]
    out = np.empty(n)
    lo, hi = boxes[:, 0, :], boxes[:, 1, :]
    for a in range(0, n, chunk):
        oo, dd = o[a : a + chunk, None, :], d[a : a + chunk, None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / dd
            t1 = (lo[None] - oo) * inv
            t2 = (hi[None] - oo) * inv
        tn = np.nanmax(np.minimum(t1, t2), axis=2)
        tf = np.nanmin(np.maximum(t1, t2), axis=2)
        hit = np.where((tf >= tn) & (tf > 0), np.where(tn > 1e-6, tn, tf), np.inf)
        hit[:, 0] = tf[:, 0]  # the hall is always hit from the inside -> far intersection
        out[a : a + chunk] = hit.min(axis=1)
    return out


# ----------------------------------------------------------------------------------------------
# LiDAR scan patterns: unit directions in the LiDAR frame (x forward, y left, z up) + time offsets
# ----------------------------------------------------------------------------------------------


def pattern_avia(n, scan_index, period):
    """Non-repetitive rosette, FOV 70.4 x 77.2 deg."""
    k = np.arange(n, dtype=np.float64)
    tau = (k / n + scan_index) * period
    a = 2 * np.pi * 97.3 * tau
    b = 2 * np.pi * 41.9 * tau + 0.37 * scan_index
    r = np.cos(2.0 * b + 0.5)
    az = np.deg2rad(35.2) * r * np.cos(a) * 0.995
    el = np.deg2rad(38.6) * r * np.sin(a) * 0.995
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=-1)
    return d, k * (period / n)


def pattern_spinning(rings, cols, vfov_deg, period, per_firing_time):
    el = np.deg2rad(np.linspace(-vfov_deg, vfov_deg, rings))
    az = -2 * np.pi * (np.arange(cols) + 0.5) / cols + np.pi  # clockwise sweep
    azg, elg = np.meshgrid(az, el, indexing="ij")  # column-major in time
    d = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], axis=-1).reshape(-1, 3)
    if per_firing_time:
        toff = np.arange(rings * cols, dtype=np.float64) * (period / (rings * cols))
    else:  # Ouster: one timestamp per column
        toff = np.repeat(np.arange(cols, dtype=np.float64) * (period / cols), rings)
    return d, toff


# ----------------------------------------------------------------------------------------------
# keyframe case
# ----------------------------------------------------------------------------------------------


@dataclass
class Shape:
    name: str
    lidar: str  # avia | ouster | vlp16
    pts_per_scan: int
    scan_period: float
    stack_interval: float
    n_features: int
    rig: Rig
    traj: Trajectory
    scene: str = "hall"
    max_range: float = 200.0


def shape(cfg: int) -> Shape:
    if cfg == 1:
        return Shape("cfg1_avia_752x480", "avia", 24000, 0.05, 0.5, 150,
                     rig_r3live((752, 480), (458.654, 457.296, 367.215, 248.375)), Trajectory(1.0, 0.2, 2.0))
    if cfg == 2:
        return Shape("cfg2_r3live_avia_1280x1024", "avia", 24000, 0.1, 0.5, 300, rig_r3live(), Trajectory(1.0, 0.2, 2.0))
    if cfg == 3:
        return Shape("cfg3_fusionportable_os1_128_1440x1080", "ouster", 131072, 0.1, 2.0, 200, rig_fusionportable(),
                     Trajectory(10.0, 0.03, 0.6), scene="street")
    if cfg == 4:
        return Shape("cfg4_vlp16_1280x1024", "vlp16", 28800, 0.1, 0.5, 500, rig_r3live(), Trajectory(1.0, 0.2, 2.0))
    raise ValueError(cfg)


@dataclass
class Keyframe:
    shape: Shape
    seed: int
    stamp: float
    points: np.ndarray  # XYZIT, time-sorted; includes points before the window and after the stamp
    ins: np.ndarray  # INS
    queries: np.ndarray  # (F, 3) f32 unit rays
    pixels: np.ndarray  # (F, 2) f32
    extra: dict = field(default_factory=dict)

    @property
    def rig(self):
        return self.shape.rig


def _scan_points(shp: Shape, boxes, rng, scan_index, scan_start):
    if shp.lidar == "avia":
        d_l, toff = pattern_avia(shp.pts_per_scan, scan_index, shp.scan_period)
    elif shp.lidar == "ouster":
        d_l, toff = pattern_spinning(128, 1024, 22.5, shp.scan_period, per_firing_time=False)
    elif shp.lidar == "vlp16":
        d_l, toff = pattern_spinning(16, 1800, 15.0, shp.scan_period, per_firing_time=True)
    else:
        raise ValueError(shp.lidar)
    t = scan_start + toff
    rig = shp.rig
    # true LiDAR pose at the time the INS would report for this point (t + td_b_l)
    R_wb, p_wb = shp.traj.pose(t + rig.td_b_l)
    R_bl = rig.R_b_c @ rig.R_c_l
    t_bl = rig.R_b_c @ rig.t_c_l + rig.t_b_c
    o_w = p_wb + R_wb @ t_bl
    d_w = np.einsum("nij,nj->ni", R_wb @ R_bl, d_l)
    rng_m = raycast(boxes, o_w, d_w)
    rng_m = rng_m + rng.normal(0.0, 0.01, size=rng_m.shape)
    # a sprinkle of returns that the range gate must drop (le_vins.cc:448-451): blind zone and > 200 m
    u = rng.random(rng_m.shape)
    rng_m = np.where(u < 0.003, rng.uniform(0.15, 0.95, rng_m.shape), rng_m)
    rng_m = np.where(u > 0.9985, rng.uniform(201.0, 240.0, rng_m.shape), rng_m)
    pts = np.zeros(t.shape[0], dtype=XYZIT)
    xyz = (d_l * rng_m[:, None]).astype(np.float32)
    pts["x"], pts["y"], pts["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    pts["intensity"] = rng.uniform(0, 255, t.shape[0]).astype(np.float32)
    pts["time"] = t
    return pts


def make_ins(shp: Shape, t_begin, t_end, rate=200.0):
    n = int(np.floor((t_end - t_begin) * rate)) + 1
    t = t_begin + np.arange(n) / rate
    R, p = shp.traj.pose(t)
    ins = np.zeros(n, dtype=INS)
    ins["time"] = t
    ins["state_time"] = t
    ins["p"] = p
    ins["q"] = matrix_to_quat(R)
    return ins


def make_queries(shp: Shape, rng, n_features=None):
    f = shp.n_features if n_features is None else n_features
    w, h = shp.rig.resolution
    px = np.stack([rng.uniform(5, w - 5, f), rng.uniform(5, h - 5, f)], axis=-1).astype(np.float32)  # tracking.cc:1103-1106
    return pixels_to_unit_rays(shp.rig.intr, px), px


def pixels_to_unit_rays(intr, px):
    """Camera::pixel2unitcam (camera.cc:123-135) then .cast<float>() (tracking.cc:391)."""
    fx, fy, cx, cy, skew = intr
    y = (px[:, 1].astype(np.float64) - cy) / fy
    x = (px[:, 0].astype(np.float64) - cx - skew * y) / fx
    v = np.stack([x, y, np.ones_like(x)], axis=-1)
    v /= np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])[:, None]
    return v.astype(np.float32)


def make_keyframe(cfg: int, seed: int | None = None, stamp_offset=3.0, n_scans_before=2, n_scans_after=1,
                  pts_per_scan=None, n_features=None, ins_margin=0.6, shp: Shape | None = None) -> Keyframe:
    """One keyframe: a time-sorted point buffer that holds `n_scans_before` scans older than the stack window,
    the window itself and `n_scans_after` scans newer than the stamp (so the window logic of
    le_vins.cc:413-437 has something to do), an INS window and the feature rays."""
    shp = shp or shape(cfg)
    if pts_per_scan is not None:
        shp = Shape(**{**shp.__dict__, "pts_per_scan": pts_per_scan})
    seed = BASE_SEED + cfg if seed is None else seed
    rng = np.random.default_rng(seed)
    boxes = make_scene(rng, shp.scene)
    # the stamp falls mid-scan so that num_projected cuts a scan in two
    stamp = T0 + stamp_offset + 0.37 * shp.scan_period
    n_win = int(round(shp.stack_interval / shp.scan_period))
    first_scan = int(np.floor((stamp - shp.stack_interval - T0) / shp.scan_period)) - n_scans_before
    last_scan = int(np.floor((stamp - T0) / shp.scan_period)) + n_scans_after
    assert last_scan - first_scan >= n_win
    if shp.lidar in ("ouster", "vlp16") and pts_per_scan is not None:
        raise ValueError("spinning patterns have a fixed point count")
    scans = [_scan_points(shp, boxes, rng, s, T0 + s * shp.scan_period) for s in range(first_scan, last_scan + 1)]
    pts = np.concatenate(scans)
    t_lo = pts["time"][0] + min(shp.rig.td_b_l, 0.0) - ins_margin
    t_hi = stamp + 0.02  # the INS window reaches just past the frame stamp (le_vins.cc:356)
    ins = make_ins(shp, t_lo, t_hi)
    q, px = make_queries(shp, rng, n_features)
    return Keyframe(shp, seed, stamp, pts, ins, q, px, {"boxes": boxes})


# ----------------------------------------------------------------------------------------------
# batch of independent keyframes over one long point stream (cfg 5 and the bench workload)
# ----------------------------------------------------------------------------------------------


@dataclass
class Batch:
    shape: Shape
    points: np.ndarray  # XYZIT stream, time-sorted
    ins: np.ndarray  # INS stream covering it
    stamps: np.ndarray  # (K,) f64
    pt_begin: np.ndarray  # (K,) i64 offset of each keyframe's buffer slice in `points`
    pt_count: np.ndarray  # (K,) i64 length of the slice (holds the window + a little slack on both sides)
    ins_begin: np.ndarray  # (K,) i64
    ins_count: np.ndarray  # (K,) i64
    queries: np.ndarray  # (K, F, 3) f32

    @property
    def n_keyframes(self):
        return len(self.stamps)


def make_batch(cfg: int, n_keyframes: int, seed: int | None = None, n_unique_scans: int | None = None,
               ins_span: float = 2.5, shp: Shape | None = None) -> Batch:
    """K non-overlapping keyframes, one every `stack_interval`, over a continuous stream.

    To bound generation time the stream repeats `n_unique_scans` ray-cast scans (timestamps keep advancing and
    every keyframe still gets its own INS poses through the moving trajectory, so every keyframe does different
    arithmetic); pass None to ray-cast every scan."""
    shp = shp or shape(cfg)
    seed = BASE_SEED + 5 if seed is None else seed
    rng = np.random.default_rng(seed)
    boxes = make_scene(rng, shp.scene)
    n_win = int(round(shp.stack_interval / shp.scan_period))
    n_scans = n_keyframes * n_win + 2
    uniq = n_scans if n_unique_scans is None else min(n_unique_scans, n_scans)
    base = [_scan_points(shp, boxes, rng, s, T0 + s * shp.scan_period) for s in range(uniq)]
    per = base[0].shape[0]
    pts = np.empty(n_scans * per, dtype=XYZIT)
    for s in range(n_scans):
        src = base[s % uniq]
        dst = pts[s * per : (s + 1) * per]
        dst[:] = src
        dst["time"] = src["time"] + (s - s % uniq) * shp.scan_period
    stamps = T0 + (np.arange(n_keyframes) + 1) * shp.stack_interval + 0.37 * shp.scan_period
    times = pts["time"]
    lo = np.searchsorted(times, stamps - shp.stack_interval - 0.3 * shp.scan_period, side="left")
    hi = np.searchsorted(times, stamps + 0.2 * shp.scan_period, side="right")
    ins = make_ins(shp, T0 + min(shp.rig.td_b_l, 0.0) - 1.0, times[-1] + 0.5)
    it = ins["time"]
    ib = np.searchsorted(it, stamps - ins_span + min(shp.rig.td_b_l, 0.0), side="left")
    ie = np.searchsorted(it, stamps + 0.02, side="right")
    ib = np.maximum(ib - 1, 0)
    q = np.stack([make_queries(shp, rng)[0] for _ in range(n_keyframes)])
    return Batch(shp, pts, ins, stamps, lo.astype(np.int64), (hi - lo).astype(np.int64), ib.astype(np.int64),
                 (ie - ib).astype(np.int64), q)


def batch_keyframe(b: Batch, k: int) -> Keyframe:
    """View keyframe k of a batch as a standalone Keyframe (for the oracle / single-keyframe API)."""
    pts = b.points[b.pt_begin[k] : b.pt_begin[k] + b.pt_count[k]]
    ins = b.ins[b.ins_begin[k] : b.ins_begin[k] + b.ins_count[k]]
    return Keyframe(b.shape, 0, float(b.stamps[k]), pts, ins, b.queries[k], np.zeros((b.queries.shape[1], 2), np.float32))
