"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a batch of keyframes, each computes its shard (the
oracle stands in for the GPU here -- this tests the partition + gather plumbing, not the kernels) and the gathered
result equals the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from levins import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for n in (1, 2, 7, 10, 256, 10000):
        for w in (1, 2, 3, 4, 8):
            r = [shard.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def _batch():
    shp = synth.Shape(**{**synth.shape(2).__dict__, "pts_per_scan": 1500, "n_features": 24})
    return synth.make_batch(2, 5, n_unique_scans=9, shp=shp)


def _compute(oracle, b):
    ty, xd = [], []
    for k in range(b.n_keyframes):
        kf = synth.batch_keyframe(b, k)
        r = oracle.keyframe(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries,
                            kf.shape.stack_interval, nthreads=1)
        ty.append(r["type"])
        xd.append(r["xyz_depth"])
    return np.concatenate(ty), np.concatenate(xd)


def _worker(rank, world, port, out_dir):
    for p in (os.path.join(ROOT, "le-vins_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import oracle_py

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = _batch()
    mine = shard.split_batch(b, world, rank)
    ty, xd = _compute(oracle_py, mine)
    gt, gx = shard.gather_results(torch.from_numpy(ty), torch.from_numpy(xd), b.n_keyframes, b.queries.shape[1], world,
                                  rank, dist=dist)
    np.save(os.path.join(out_dir, f"type_{rank}.npy"), gt.numpy())
    np.save(os.path.join(out_dir, f"xyzd_{rank}.npy"), gx.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_single_process(oracle, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b = _batch()
    ty, xd = _compute(oracle, b)
    assert (ty == 1).sum() > 0
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"type_{r}.npy"), ty)
        assert np.array_equal(np.load(tmp_path / f"xyzd_{r}.npy"), xd)


def test_split_batch_is_self_contained():
    b = _batch()
    parts = [shard.split_batch(b, 2, r) for r in range(2)]
    assert sum(p.n_keyframes for p in parts) == b.n_keyframes
    k = 0
    for p in parts:
        for j in range(p.n_keyframes):
            a, c = synth.batch_keyframe(p, j), synth.batch_keyframe(b, k)
            assert a.stamp == c.stamp and np.array_equal(a.points, c.points) and np.array_equal(a.ins, c.ins)
            assert np.array_equal(a.queries, c.queries)
            k += 1
    with pytest.raises(ValueError):
        shard.split_batch(b, 8, 7)


def _pipe_worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "le-vins_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 17 * 40
    local = torch.zeros(n, dtype=torch.uint8)
    pipe = shard.PipelinedGather(torch, dist, local, world, torch.device("cpu"))
    seen = []
    for step in range(5):  # the "library" overwrites its block every step; each gather must see that step's content
        local.fill_(10 * step + rank)
        j = pipe.submit()
        local.fill_(255)  # overwritten right away, like the next batch_run would
        if step >= 1:
            pipe.pending[j ^ 1].wait()
            seen.append(pipe.gathered[j ^ 1].clone())
    pipe.drain()
    seen.append(pipe.latest().clone())
    ok = all(bool((g.view(world, n)[r] == 10 * s + r).all()) for s, g in enumerate(seen) for r in range(world))
    np.save(os.path.join(out_dir, f"pipe_{rank}.npy"), np.array([ok, len(seen)]))
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_gather_two_ranks(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_pipe_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        ok, n = np.load(tmp_path / f"pipe_{r}.npy")
        assert ok == 1 and n == 5
