"""The packed host format of levins_batch_upload_packed (16-byte records: x, y, z, u32 nanoseconds after the scan's stamp):
abi.pack_points predicts the PointXYZIT records the device reconstructs; checked here on the CPU (layout, losslessness of the
coordinates, times within half a nanosecond and identical to the formula of ROS/lidar/lidar_converter.cc:162)."""
import numpy as np

from levins import abi, synth


def test_pack_points_layout_and_times():
    assert abi.PACKED.itemsize == 16 and abi.SCAN_DESC.itemsize == 24
    shp = synth.Shape(**{**synth.shape(3).__dict__, "stack_interval": 0.2})
    b = synth.make_batch(3, 2, n_unique_scans=3, shp=shp, ins_span=1.5)
    per = b.shape.pts_per_scan
    first = np.arange(len(b.points) // per) * per
    packed, scans, exact = abi.pack_points(b.points, first, b.points["time"][::per])
    assert len(packed) == len(exact) == len(b.points) and len(scans) == len(first)
    assert np.array_equal(scans["first"], first) and np.all(scans["count"] == per)
    for c in "xyz":
        assert packed[c].tobytes() == b.points[c].tobytes() == exact[c].tobytes()  # coordinates are carried as they are
    assert np.max(np.abs(exact["time"] - b.points["time"])) <= 0.5e-9 + 2 * np.spacing(345600.0)  # half a ns + f64 rounding
    stamp = np.repeat(scans["stamp"], scans["count"].astype(np.int64))
    assert np.array_equal(exact["time"], stamp + packed["dt_ns"].astype(np.float64) * 1.0e-9)  # the device's formula
    assert np.all(np.diff(exact["time"]) >= 0)  # still time-sorted
    # packing what the device reconstructs is a fixed point
    packed2, _, exact2 = abi.pack_points(exact, first, exact["time"][::per])
    assert packed2.tobytes() == packed.tobytes() and exact2.tobytes() == exact.tobytes()
