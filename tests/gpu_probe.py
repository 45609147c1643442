"""Ad-hoc GPU diagnostics (not collected by pytest)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "le-vins_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from levins import abi, synth
import oracle_py as orc

b = synth.make_batch(2, 6, n_unique_scans=7, shp=synth.Shape(**{**synth.shape(2).__dict__, "pts_per_scan": 8000}))
ctx = abi.Context(0)
ctx.set_extrinsics(b.shape.rig.pose_b_c, b.shape.rig.pose_c_l, b.shape.rig.td_b_l)
ctx.batch_set_wave(4)
ctx.batch_upload(b.points, b.ins, b.queries.reshape(-1, 3), abi.batch_descs(b))
outs = []
for r in range(4):
    ctx.batch_run()
    outs.append(ctx.batch_download())
for r in range(1, 4):
    for k in outs[0]:
        d = outs[0][k] != outs[r][k]
        if d.any():
            idx = np.argwhere(d)[:5]
            print("run", r, "key", k, "ndiff", int(d.sum()), "first", idx.tolist(), outs[0][k][tuple(idx[0])], outs[r][k][tuple(idx[0])])
f = b.queries.shape[1]
for k in range(b.n_keyframes):
    kf = synth.batch_keyframe(b, k)
    ref = orc.keyframe(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries, kf.shape.stack_interval)
    for r in range(2):
        ty = outs[r]["type"][k*f:(k+1)*f]
        print("kf", k, "run", r, "nproj", outs[r]["n_projected"][k], ref["n_projected"], "m", outs[r]["m"][k], ref["m"], "type agree", np.mean(ty == ref["type"]))
