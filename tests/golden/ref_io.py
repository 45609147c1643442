"""Exchange files between the committed golden fixtures and oracle/ref_harness/ref_golden (the UNMODIFIED reference compiled
on a box that has its dependencies; see oracle/ref_harness/BUILD.md).

    python tests/golden/ref_io.py dump <dir>    # <dir>/<case>/{points,ins,queries,meta}.bin from tests/golden/<case>.npz
    python tests/golden/ref_io.py pack <dir>    # tests/golden/ref_<case>.npz from <dir>/<case>/ref_*.bin
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
XYZI = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"), ("intensity", "<f4"), ("pad", "<f4", (3,))])


def cases():
    return sorted(p for p in glob.glob(os.path.join(HERE, "*.npz")) if not os.path.basename(p).startswith("ref_"))


def dump(out):
    for path in cases():
        g = np.load(path)
        d = os.path.join(out, os.path.basename(path)[:-4])
        os.makedirs(d, exist_ok=True)
        g["points"].tofile(os.path.join(d, "points.bin"))
        ins = g["ins"]
        np.concatenate([ins["time"][:, None], ins["state_time"][:, None], ins["p"], ins["q"]], axis=1).astype(np.float64).tofile(
            os.path.join(d, "ins.bin"))
        g["queries"].astype(np.float32).tofile(os.path.join(d, "queries.bin"))
        meta = np.concatenate([[float(g["stamp"]), float(g["stack_interval"]), float(g["td_b_l"])], g["pose_b_c"], g["pose_c_l"],
                               [1000.0, 1000.0, 720.0, 540.0]])  # intrinsics are not used by stages 1-3
        meta.astype(np.float64).tofile(os.path.join(d, "meta.bin"))
        print("dumped", d)


def pack(out):
    for path in cases():
        name = os.path.basename(path)[:-4]
        d = os.path.join(out, name)
        if not os.path.exists(os.path.join(d, "ref_summary.txt")):
            print("no reference outputs for", name)
            continue
        summ = dict(line.split() for line in open(os.path.join(d, "ref_summary.txt")).read().strip().split("\n"))
        f = len(np.load(path)["queries"])
        unit = os.path.join(d, "ref_unit.bin")
        np.savez_compressed(
            os.path.join(HERE, f"ref_{name}.npz"), first=np.int64(summ["first"]), n=np.int64(summ["n"]), m=np.int64(summ["m"]),
            cloud=np.fromfile(os.path.join(d, "ref_cloud.bin"), XYZI),
            unit_sphere=np.fromfile(unit, XYZI) if os.path.exists(unit) else np.zeros(0, XYZI),
            type=np.fromfile(os.path.join(d, "ref_type.bin"), np.int8),
            xyz_depth=np.fromfile(os.path.join(d, "ref_xyzd.bin"), np.float32).reshape(f, 4),
            knn_idx=np.fromfile(os.path.join(d, "ref_knn_idx.bin"), np.int32).reshape(f, 5),
            knn_d2=np.fromfile(os.path.join(d, "ref_knn_d2.bin"), np.float32).reshape(f, 5))
        print("packed", name)


if __name__ == "__main__":
    {"dump": dump, "pack": pack}[sys.argv[1]](sys.argv[2])
