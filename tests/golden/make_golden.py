"""Generate the committed golden fixtures: small seeded inputs plus the ORACLE's outputs at every stage boundary.

The reference itself has no golden vectors and cannot be built here, so these pin the oracle's
behaviour (any later change of the restatement shows up as a diff) and give the GPU path a fixture that does not
need the oracle at all.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "le-vins_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle_py as orc  # noqa: E402
from levins import synth  # noqa: E402

CASES = {
    "avia_small": dict(cfg=2, pts_per_scan=3000, n_features=64),
    "vlp16_sparse": dict(cfg=4, n_features=96, decimate=4),
    "ouster_td": dict(cfg=3, n_features=64),  # front 90-degree sector, every 2nd column, 0.2 s stack (td_b_l = -0.48)
}


def build(name, spec):
    spec = dict(spec)
    dec = spec.pop("decimate", None)
    if spec["cfg"] == 3:
        shp = synth.Shape(**{**synth.shape(3).__dict__, "stack_interval": 0.2})
        kf = synth.make_keyframe(3, shp=shp, n_features=spec["n_features"], n_scans_before=1)
        p = kf.points
        col = (np.arange(len(p)) // 128) % 2 == 0
        kf.points = p[(p["x"] > 0) & (np.abs(p["y"]) < p["x"]) & col]
    else:
        kf = synth.make_keyframe(**spec)
    pts = kf.points[::dec] if dec else kf.points
    interval = kf.shape.stack_interval
    s1 = orc.stack_deskew(pts, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, interval, details=True)
    s2 = orc.range_image(s1["cloud"])
    s3 = orc.associate(s2["unit_sphere"], kf.queries)
    return dict(
        points=pts, ins=kf.ins, queries=kf.queries, stamp=np.float64(kf.stamp), stack_interval=np.float64(interval),
        pose_b_c=kf.rig.pose_b_c, pose_c_l=kf.rig.pose_c_l, td_b_l=np.float64(kf.rig.td_b_l),
        first=np.int64(s1["first"]), n=np.int64(s1["n"]), cloud=s1["cloud"], kept=s1["kept"], ins_index=s1["ins_index"],
        winner=s2["winner"], unit_sphere=s2["unit_sphere"], m=np.int64(s2["m"]),
        type=s3["type"], xyz_depth=s3["xyz_depth"], knn_idx=s3["knn_idx"], knn_d2=s3["knn_d2"], flags=s3["flags"],
    )


def main():
    for name, spec in CASES.items():
        d = build(name, spec)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **d)
        print(name, "points", len(d["points"]), "n", int(d["n"]), "M", int(d["m"]), "associated", int((d["type"] == 1).sum()),
              f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
