"""Driver record formats of the reference's LidarConverter (ROS/lidar/lidar_converter.cc, le_vins/lidar/lidar.h:36-76) as
levins_raw_format descriptors, and synthetic wire records for tests and bench (harness side only)."""
from __future__ import annotations

import numpy as np

RAW_FORMAT = np.dtype([(k, "<u4") for k in ("kind", "point_step", "off_x", "off_y", "off_z", "off_intensity", "off_time",
                                             "off_ring", "off_tag", "time_kind", "ring_bytes", "pad")])

# in-memory livox_ros_driver::CustomPoint (u32 offset_time, f32 x y z, u8 reflectivity tag line, 1 B tail padding)
LIVOX = np.dtype([("offset_time", "<u4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("reflectivity", "u1"), ("tag", "u1"),
                  ("line", "u1"), ("pad", "u1")])
# the pcl point structs the reference unpacks PointCloud2 into (lidar.h:36-76; EIGEN_ALIGN16)
VELODYNE = np.dtype({"names": ["x", "y", "z", "intensity", "time", "ring"], "formats": ["<f4", "<f4", "<f4", "<f4", "<f4", "<u2"],
                     "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
OUSTER = np.dtype({"names": ["x", "y", "z", "intensity", "t", "reflectivity", "ambient", "range", "ring"],
                   "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "<u2", "<u4", "u1"],
                   "offsets": [0, 4, 8, 16, 20, 24, 26, 28, 32], "itemsize": 48})
HESAI = np.dtype({"names": ["x", "y", "z", "intensity", "timestamp", "ring"], "formats": ["<f4", "<f4", "<f4", "<f4", "<f8", "<u2"],
                  "offsets": [0, 4, 8, 16, 24, 32], "itemsize": 48})


def _fmt(**kw):
    f = np.zeros(1, RAW_FORMAT)
    for k, v in kw.items():
        f[k] = v
    return f


def fmt_livox(point_step=LIVOX.itemsize):
    return _fmt(kind=1, point_step=point_step, off_time=0, off_x=4, off_y=8, off_z=12, off_intensity=16, off_tag=17, off_ring=18)


def fmt_livox_wire():
    """The 19-byte serialized ROS record (no tail padding): same offsets, unaligned stride."""
    return fmt_livox(19)


def fmt_velodyne():
    return _fmt(kind=2, point_step=32, off_x=0, off_y=4, off_z=8, off_intensity=16, off_time=20, off_ring=24, time_kind=1, ring_bytes=2)


def fmt_ouster():
    return _fmt(kind=2, point_step=48, off_x=0, off_y=4, off_z=8, off_intensity=16, off_time=20, off_ring=32, time_kind=2, ring_bytes=1)


def fmt_hesai():
    return _fmt(kind=2, point_step=48, off_x=0, off_y=4, off_z=8, off_intensity=16, off_time=24, off_ring=32, time_kind=3, ring_bytes=2)


GPS_UNIX_OFFSET = 315964800 - 18  # gpstime.h:33


def livox_records(points, week: int, timebase_ns: int | None = None, rng=None, bad_fraction=0.0):
    """PointXYZIT (GPS seconds of week) -> CustomPoint records whose conversion gives the same xyz; returns (records,
    timebase_ns).  `bad_fraction` of the records get a tag / line the converter drops."""
    unix = points["time"] + week * 604800 + GPS_UNIX_OFFSET
    if timebase_ns is None:
        timebase_ns = int(np.floor(unix.min() * 1e9))
    rec = np.zeros(len(points), LIVOX)
    rec["offset_time"] = np.clip(np.round(unix * 1e9 - timebase_ns), 0, 2**32 - 1).astype(np.uint32)
    rec["x"], rec["y"], rec["z"] = points["x"], points["y"], points["z"]
    rec["reflectivity"] = np.clip(points["intensity"], 0, 255).astype(np.uint8)
    rec["tag"] = 0x10 if rng is None else rng.choice([0x00, 0x10, 0x01, 0x12], len(points))
    rec["line"] = 0 if rng is None else rng.integers(0, 6, len(points))
    if rng is not None and bad_fraction > 0:
        bad = rng.random(len(points)) < bad_fraction
        rec["tag"][bad] = rng.choice([0x20, 0x30, 0x31], bad.sum())
        bad2 = rng.random(len(points)) < bad_fraction
        rec["line"][bad2] = rng.integers(6, 9, bad2.sum())
    return rec, timebase_ns


def as_bytes(records, point_step=None):
    """Structured records -> flat uint8, optionally re-strided (e.g. 20 -> 19 for the serialized Livox record)."""
    b = np.ascontiguousarray(records).view(np.uint8).reshape(len(records), records.dtype.itemsize)
    if point_step is not None and point_step != records.dtype.itemsize:
        b = np.ascontiguousarray(b[:, :point_step])
    return b.reshape(-1)
