"""Multi-GPU host logic: independent keyframes are sharded across ranks, results are gathered once at the end.

The path has no data dependency between keyframes, so there is no data-path collective; the only
exchange step is the final gather of the per-feature results (type i8 + x, y, z, depth f32 = 17 B / feature), issued
with torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition (keeps the scans shared by neighbouring keyframes on one rank)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_results(local_type, local_xyzd, n_items_total: int, features_per_item: int, world: int, rank: int, dist=None,
                   device=None):
    """All-gather per-feature results of a block-sharded batch.  `local_*` are torch tensors (on `device` for NCCL).
    Returns (type[n_total*F], xyz_depth[n_total*F, 4]) in keyframe order on every rank."""
    import torch

    if world == 1:
        return local_type, local_xyzd
    f = features_per_item
    counts = [shard_range(n_items_total, world, r)[1] - shard_range(n_items_total, world, r)[0] for r in range(world)]
    pad = max(counts) * f
    t = torch.full((pad,), -1, dtype=torch.int8, device=device)
    x = torch.zeros((pad, 4), dtype=torch.float32, device=device)
    t[: local_type.numel()] = local_type
    x[: local_xyzd.shape[0]] = local_xyzd
    gt = torch.empty((world * pad,), dtype=torch.int8, device=device)
    gx = torch.empty((world * pad, 4), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(gt, t)
    dist.all_gather_into_tensor(gx, x)
    ts, xs = [], []
    for r in range(world):
        ts.append(gt[r * pad: r * pad + counts[r] * f])
        xs.append(gx[r * pad: r * pad + counts[r] * f])
    return torch.cat(ts), torch.cat(xs)


class PipelinedGather:
    """The final gather of a stream of batches, overlapped with the next batch's kernels.

    `local` is the rank's contiguous result block in library memory (xyz_depth then type, 17 B / feature,
    levins_batch_result_pointers).  submit() snapshots it into one of two staging buffers (device-to-device on the compute
    stream) and starts an asynchronous all-gather from there, so the library may overwrite its block while the collective
    is still in flight; the gather issued two submits ago is waited for before its buffers are reused.  drain() waits for
    everything outstanding (call it before reading `gathered` or stopping a clock)."""

    def __init__(self, torch, dist, local, world: int, device):
        self.torch, self.dist, self.local = torch, dist, local
        n = local.numel()
        self.stage = [torch.empty((n,), dtype=local.dtype, device=device) for _ in range(2)]
        self.gathered = [torch.empty((world * n,), dtype=local.dtype, device=device) for _ in range(2)]
        self.pending = [None, None]
        self.k = 0

    def submit(self):
        j = self.k & 1
        self.k += 1
        if self.pending[j] is not None:
            self.pending[j].wait()
        self.stage[j].copy_(self.local, non_blocking=True)
        self.pending[j] = self.dist.all_gather_into_tensor(self.gathered[j], self.stage[j], async_op=True)
        return j

    def drain(self):
        for j in range(2):
            if self.pending[j] is not None:
                self.pending[j].wait()
                self.pending[j] = None

    def latest(self):
        """The most recently gathered block (after drain())."""
        return self.gathered[(self.k - 1) & 1]


def split_batch(batch, world: int, rank: int):
    """The sub-batch of a synth.Batch that `rank` owns (descriptors re-based onto the rank's own pools)."""
    from . import synth

    b, e = shard_range(batch.n_keyframes, world, rank)
    if e == b:
        raise ValueError("more ranks than keyframes")
    p0, p1 = int(batch.pt_begin[b]), int((batch.pt_begin[b:e] + batch.pt_count[b:e]).max())
    i0, i1 = int(batch.ins_begin[b:e].min()), int((batch.ins_begin[b:e] + batch.ins_count[b:e]).max())
    return synth.Batch(batch.shape, batch.points[p0:p1], batch.ins[i0:i1], batch.stamps[b:e], batch.pt_begin[b:e] - p0,
                       batch.pt_count[b:e], batch.ins_begin[b:e] - i0, batch.ins_count[b:e], batch.queries[b:e])
