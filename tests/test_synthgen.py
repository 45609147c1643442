"""The config-5 workload generator (harness): the device-side generator and its numpy twin produce the same bits, and a batch
generated on the device and attached without any copy (levins_batch_attach_device) gives the oracle's results."""
import numpy as np
import pytest

from levins import synth, synthgen


def test_numpy_twin_is_deterministic_and_sane(oracle):
    sp = synthgen.Spec(scans_per_kf=1, n_feat=64)
    a = synthgen.keyframe_numpy(sp, 5)
    b = synthgen.keyframe_numpy(sp, 5)
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))
    pts, ins, q = a
    assert len(pts) == synthgen.PER_SCAN and np.all(np.diff(pts["time"]) >= 0) and np.all(np.diff(ins["time"]) > 0)
    assert np.allclose(np.linalg.norm(q.astype(np.float64), axis=1), 1.0, atol=1e-6)
    assert np.allclose(np.linalg.norm(ins["q"], axis=1), 1.0, atol=1e-12)
    other = synthgen.keyframe_numpy(sp, 6)[0]
    assert other.tobytes() != pts.tobytes()  # per-keyframe seeds
    ref = oracle.keyframe(pts, sp.stamp(5), ins, sp.rig.pose_b_c, sp.rig.pose_c_l, sp.td_b_l, q, sp.stack_interval)
    assert ref["m"] > 5000 and (ref["type"] == 1).sum() > 10


@pytest.mark.gpu
def test_device_generator_equals_numpy_twin_and_batch_equals_oracle(oracle, built_lib):
    import torch

    from levins import abi

    sp = synthgen.Spec(scans_per_kf=5, n_feat=120)
    gen = synthgen.DeviceGenerator(torch, 0, sp)
    first, n_kf = 40, 6
    d_pts, d_ins, d_q = gen.generate(first, n_kf)
    torch.cuda.synchronize()
    h_pts = d_pts.cpu().numpy().view(synth.XYZIT).reshape(n_kf, sp.per_kf)
    h_ins = d_ins.cpu().numpy().view(synth.INS).reshape(n_kf, sp.n_ins)
    h_q = d_q.cpu().numpy().reshape(n_kf, sp.n_feat, 3)
    dirs = synthgen.directions()
    twins = [synthgen.keyframe_numpy(sp, first + k, dirs) for k in range(n_kf)]
    for k, (pts, ins, q) in enumerate(twins):
        assert h_pts[k].tobytes() == pts.tobytes(), f"keyframe {k}: points differ"
        assert h_ins[k].tobytes() == ins.tobytes(), f"keyframe {k}: INS differs"
        assert h_q[k].tobytes() == q.tobytes(), f"keyframe {k}: queries differ"
    c = abi.Context(0, stack_interval=sp.stack_interval)
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    c.set_extrinsics(sp.rig.pose_b_c, sp.rig.pose_c_l, sp.td_b_l)
    c.batch_set_wave(4)
    c.batch_attach_device(d_pts.data_ptr(), n_kf * sp.per_kf, d_ins.data_ptr(), n_kf * sp.n_ins, d_q.data_ptr(), n_kf * sp.n_feat,
                          gen.descs(first, n_kf))
    c.batch_run()
    got = c.batch_download()
    f = sp.n_feat
    for k, (pts, ins, q) in enumerate(twins):
        ref = oracle.keyframe(pts, sp.stamp(first + k), ins, sp.rig.pose_b_c, sp.rig.pose_c_l, sp.td_b_l, q, sp.stack_interval)
        assert got["n_projected"][k] == ref["n_projected"] and got["m"][k] == ref["m"]
        assert np.array_equal(got["type"][k * f:(k + 1) * f], ref["type"])
        assert got["xyz_depth"][k * f:(k + 1) * f].tobytes() == ref["xyz_depth"].tobytes()
    assert (got["type"] == 1).sum() > 100
    c.close()
