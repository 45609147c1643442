"""Run under torchrun with N >= 2 ranks (one per GPU): a seeded batch is block-sharded over the ranks, every rank runs its
shard through the C ABI, the results are exchanged both ways the ABI offers (peer mode: k_plane writes straight into every
rank's gathered block over NVLink; NCCL: levins_gather), and every rank checks that the gathered block equals, BYTE FOR BYTE,
what a single GPU computes for the whole batch.  Prints MULTI_GPU_OK <world> on rank 0.  Used by tests/test_multi_gpu.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "le-vins_b200"))


def main():
    import torch
    import torch.distributed as dist

    from levins import abi, shard, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    per_rank = 3
    shp = synth.Shape(**{**synth.shape(2).__dict__, "pts_per_scan": 6000, "n_features": 97})  # 3 x 97 queries per rank: not a multiple of 16
    whole = synth.make_batch(2, per_rank * world, n_unique_scans=9, shp=shp)  # same seed on every rank
    F = whole.queries.shape[1]
    mine = shard.split_batch(whole, world, rank)
    nq = mine.n_keyframes * F

    # single-GPU answer for the whole batch (every rank computes it on its own GPU)
    ref_ctx = abi.Context(local)
    ref_ctx.set_extrinsics(whole.shape.rig.pose_b_c, whole.shape.rig.pose_c_l, whole.shape.rig.td_b_l)
    ref_ctx.batch_upload(whole.points, whole.ins, whole.queries.reshape(-1, 3), abi.batch_descs(whole))
    ref_ctx.batch_run()
    ref = ref_ctx.batch_download()
    ref_ctx.close()
    want_t = ref["type"].reshape(world, nq)
    want_x = ref["xyz_depth"].reshape(world, nq, 4)

    def slots(g):  # the gathered slots are padded to a multiple of 16 queries
        return g["type"][:, :nq], g["xyz_depth"][:, :nq]

    assert (want_t == 1).sum() > 20

    ctx = abi.Context(local)
    ctx.set_extrinsics(mine.shape.rig.pose_b_c, mine.shape.rig.pose_c_l, mine.shape.rig.td_b_l)
    ctx.batch_upload(mine.points, mine.ins, mine.queries.reshape(-1, 3), abi.batch_descs(mine))

    # (1) NCCL all-gather through the ABI
    uid = torch.from_numpy(abi.Context.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
    dist.broadcast(uid, 0)
    comm = ctx.comm_init(uid.cpu().numpy(), world, rank)
    ctx.batch_run()
    ctx.gather(comm, world)
    gt, gx = slots(ctx.gathered_download(world))
    assert gt.tobytes() == want_t.tobytes(), f"rank {rank}: NCCL-gathered types differ from the single-GPU run"
    assert gx.tobytes() == want_x.tobytes(), f"rank {rank}: NCCL-gathered depths differ"
    abi.Context.comm_destroy(comm)

    # (2) peer mode: fused into k_plane, three steps so that both halves of the double buffer are used
    h = torch.from_numpy(ctx.peer_export(world, nq)).to(dev)
    allh = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(allh, h)
    ctx.peer_attach(rank, world, torch.stack(allh).cpu().numpy())
    dist.barrier()
    for step in range(3):
        ctx.batch_run()
        ctx.peer_wait()
        gt, gx = slots(ctx.gathered_download(world))
        assert gt.tobytes() == want_t.tobytes(), f"rank {rank} step {step}: peer-gathered types differ"
        assert gx.tobytes() == want_x.tobytes(), f"rank {rank} step {step}: peer-gathered depths differ"
    dist.barrier()
    ctx.peer_detach()
    dist.barrier()
    ctx.close()
    ok = torch.ones(1, device=dev)
    dist.all_reduce(ok)
    if rank == 0:
        print(f"MULTI_GPU_OK {world} ranks={int(ok.item())} associated={(want_t == 1).sum()}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
