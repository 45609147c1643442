#!/usr/bin/env python
"""bench.py -- LiDAR depth-enhancement hot path on B200: features depth-associated / s, LiDAR pts / s, p50 keyframe ms.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one pass of the whole hot path (window search + INS table, per-point deskew fused with range-image
binning, per-feature 5-NN + plane fit) over one batch of independent synthetic keyframes of BASELINE.json configs[2],
the largest single-GPU configuration (FusionPortableV2 vehicle shape: Ouster OS1-128, 131 072 pts/scan with per-point
timestamps, 20-scan = 2.0 s stack window = 2.62 M points per keyframe, td_b_l = -0.48 s, 1440x1080 camera, 200 features
per keyframe).  Every rank processes its own batch (weak scaling); keyframes are independent, so the only exchange step
is the final gather of the per-feature results.

Printed JSON (one line, rank 0):
  value      whole-job features/s with the batch already resident in HBM (CUDA events on the launching stream)
  e2e        the same metric through the C ABI with HOST buffers: host->device copies of points / INS / queries and
             the device->host read of the results are inside the timed region
  roofline   dominant kernel (k_deskew_bin): algorithmic bytes (32 B per projected point) / its mean launch time
             measured live with CUDA events, against MEASURED_PEAKS.json's hbm_gbs
  cpu_baseline  the CPU oracle (a port of the reference's algorithm with its parallel structure) on a bounded sample
  parity     the GPU results of the timed batch against the oracle's on the keyframes the cpu_baseline leg computed
`--impl reference` times that CPU port alone on the box's host cores (the reference itself cannot be built here).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "le-vins_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "features_depth_associated_per_s"
UNIT = "features/s"
CFG = 3  # BASELINE.json configs[2]: the largest single-GPU configuration (the metric names no configuration)
DEFAULT_KEYFRAMES = 32
UNIQUE_SCANS = 20  # one stack window of ray-cast scans, repeated with advancing timestamps (every keyframe has its own poses)
INS_SPAN = 3.5  # seconds of INS before the stamp: the 2.0 s window shifted by td_b_l = -0.48 s
STREAM_CFG = 2  # p50 keyframe latency: BASELINE.json configs[1], the streaming single-GPU shape


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def workload_name(n_kf):
    return f"cfg3_fusionportable_os1_128_x20scans_1440x1080_F200_batch{n_kf}_independent_keyframes"


def cached_batch(synth, key, make):
    """Synthetic batches are deterministic in their arguments; LEVINS_BATCH_CACHE=<dir> keeps them on disk so that several
    bench.py processes on one box (A/B timing of kernel variants, the two arms) ray-cast each batch once."""
    cache = os.environ.get("LEVINS_BATCH_CACHE")
    if not cache:
        return make()
    os.makedirs(cache, exist_ok=True)
    path = os.path.join(cache, key + ".npz")
    fields = ("points", "ins", "stamps", "pt_begin", "pt_count", "ins_begin", "ins_count", "queries")
    if os.path.exists(path):
        z = np.load(path)
        return synth.Batch(synth.shape(int(z["cfg"])), *[z[f] for f in fields])
    b = make()
    tmp = path + f".{os.getpid()}.tmp.npz"
    np.savez(tmp, cfg=np.int64(int(key.split("_")[0][3:])), **{f: getattr(b, f) for f in fields})
    os.replace(tmp, path)
    return b


def make_workload(synth, n_kf, rank=0):
    seed = synth.BASE_SEED + 3 + 1000 * rank
    return cached_batch(synth, f"cfg{CFG}_k{n_kf}_s{seed}_u{UNIQUE_SCANS}",
                        lambda: synth.make_batch(CFG, n_kf, seed=seed, n_unique_scans=UNIQUE_SCANS, ins_span=INS_SPAN))


def workload_config(batch, n_kf):
    """The `config` object of the JSON line: identical keys and values in both arms."""
    return {"workload": workload_name(n_kf), "keyframes_per_gpu": int(n_kf),
            "scans_per_keyframe": int(round(batch.shape.stack_interval / batch.shape.scan_period)),
            "points_per_scan": int(batch.shape.pts_per_scan), "features_per_keyframe": int(batch.queries.shape[1]),
            "stack_interval_s": float(batch.shape.stack_interval), "td_b_l_s": float(batch.shape.rig.td_b_l),
            "resolution": list(batch.shape.rig.resolution),
            "l2": "inputs larger than L2, no flush: %d MB of points per GPU per step against 126 MB of L2" % (
                int(n_kf) * int(round(batch.shape.stack_interval / batch.shape.scan_period)) * int(batch.shape.pts_per_scan) * 32 // 1000000)}


# ------------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# CPU legs (oracle port)
# ------------------------------------------------------------------------------------------------------------
def host_cores():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU legs pass the count explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_keyframes(batch, ks, nthreads, budget_s=None, results=None):
    """Oracle port on keyframes `ks` (stopping after `budget_s` seconds of CPU work when given).  Returns
    (seconds, features, points, keyframes); `results` (a dict) receives the oracle's result per keyframe index."""
    import oracle_py as orc
    from levins import synth

    t0 = time.perf_counter()
    nf = npts = nk = 0
    for k in ks:
        kf = synth.batch_keyframe(batch, k)
        r = orc.keyframe(kf.points, kf.stamp, kf.ins, kf.rig.pose_b_c, kf.rig.pose_c_l, kf.rig.td_b_l, kf.queries,
                         kf.shape.stack_interval, nthreads=nthreads)
        if results is not None:
            results[k] = r
        nf += len(kf.queries)
        npts += r["n_projected"]
        nk += 1
        if budget_s is not None and time.perf_counter() - t0 >= budget_s:
            break
    return time.perf_counter() - t0, nf, npts, nk


def parity_report(gpu, oracle_results, F):
    """GPU results of the timed batch against the oracle's, keyframe by keyframe (bars: north_star)."""
    if not oracle_results:
        return None
    ks = sorted(oracle_results)
    type_eq = depth_ok = depth_n = depth_same = 0
    m_eq = np_eq = 0
    max_rel = 0.0
    for k in ks:
        r = oracle_results[k]
        gt, gx = gpu["type"][k * F:(k + 1) * F], gpu["xyz_depth"][k * F:(k + 1) * F]
        type_eq += int(np.sum(gt == r["type"]))
        both = (gt == 1) & (r["type"] == 1)
        d0 = r["xyz_depth"][both, 3].astype(np.float64)
        d1 = gx[both, 3].astype(np.float64)
        rel = np.abs(d1 - d0) / np.abs(d0) if both.any() else np.zeros(0)
        depth_ok += int(np.sum(rel <= 1e-4))
        depth_same += int(np.sum(d0 == d1))
        depth_n += int(both.sum())
        max_rel = max(max_rel, float(rel.max()) if rel.size else 0.0)
        m_eq += int(gpu["m"][k] == r["m"])
        np_eq += int(gpu["n_projected"][k] == r["n_projected"])
    n = len(ks)
    return {"keyframes_checked": n, "features_checked": n * F, "type_agreement": type_eq / (n * F),
            "depth_within_1e-4": (depth_ok / depth_n) if depth_n else 1.0,
            "depth_bit_identical": (depth_same / depth_n) if depth_n else 1.0, "depth_compared": depth_n,
            "max_depth_rel_err": max_rel, "m_equal": m_eq / n, "n_projected_equal": np_eq / n,
            "against": "oracle port (oracle/levins_oracle.c) on the same keyframes of the timed batch"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port: stage 1 parallel over all host threads, stages
    2-3 serial with a KD-tree, as in le_vins.cc:464 / tracking.cc:307-362,399-413) on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_py as orc
    from levins import synth

    orc.build()
    n_sample = 4  # keyframes per step (~0.2 s of host work each on 8+ threads): the default 20 steps end well within a minute
    batch = make_workload(synth, n_sample)
    cores = host_cores()
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_keyframes(batch, range(n_sample), cores)
    t_total, nf_total, np_total = 0.0, 0, 0
    for _ in range(args.steps):
        t, nf, npts, _ = cpu_keyframes(batch, range(n_sample), cores)
        t_total += t
        nf_total += nf
        np_total += npts
    value = nf_total / t_total
    F = batch.queries.shape[1]
    sample = (f"{n_sample} keyframes of the workload per step ({n_sample * F} features, {np_total // args.steps} points); "
              f"stage 1 on {cores} threads, stages 2-3 serial + KD-tree like the reference")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(batch, args.keyframes),
        "lidar_pts_per_s": np_total / t_total, "keyframes_per_s": n_sample * args.steps / t_total,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port of the reference's CPU path; the reference itself needs Eigen/PCL/FLANN/TBB/OpenCV (absent)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def pinned_copy(torch, arr):
    """numpy array -> pinned host memory (numpy view over a pinned torch tensor)."""
    raw = arr.view(np.uint8).reshape(-1)
    t = torch.empty(raw.shape[0], dtype=torch.uint8, pin_memory=True)
    v = t.numpy()
    v[:] = raw
    return t, v.view(arr.dtype).reshape(arr.shape)


def bind_near_gpu(index):
    """Pin this process to the cores of the GPU's NUMA node, so that the page-locked host pools it allocates (first touch) and
    the thread that feeds the copies are local to the GPU's PCIe root.  Does nothing when the topology is not exposed."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


def world_from_env(args):
    """World size comes from the launcher (torchrun exports WORLD_SIZE); --gpus must agree with it."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}: launch N ranks with "
                         f"`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N`")
    return world


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from levins import abi, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = world_from_env(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    n_kf = args.keyframes
    bind_near_gpu(local_rank)  # before the page-locked pools are allocated (first touch decides their NUMA node)
    batch = make_workload(synth, n_kf, rank)
    descs = abi.batch_descs(batch)
    F = batch.queries.shape[1]
    q_flat = np.ascontiguousarray(batch.queries.reshape(-1, 3))
    # The host-side form of the points: packed 16-byte records (x, y, z, u32 nanoseconds after the scan's stamp -- the driver's
    # own time field).  batch.points becomes what the device reconstructs from them (times = stamp + dt * 1e-9), so the
    # resident measurement, the e2e measurement and the CPU oracle all see the same values.
    per = batch.shape.pts_per_scan
    packed_np, scans, batch.points = abi.pack_points(batch.points, np.arange(len(batch.points) // per) * per, batch.points["time"][::per])
    keep = []
    t_pts, pts = pinned_copy(torch, packed_np)
    del packed_np
    t_ins, ins = pinned_copy(torch, batch.ins)
    t_q, q = pinned_copy(torch, q_flat)
    keep += [t_pts, t_ins, t_q]

    ctx = abi.Context(local_rank, stack_interval=batch.shape.stack_interval)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_extrinsics(batch.shape.rig.pose_b_c, batch.shape.rig.pose_c_l, batch.shape.rig.td_b_l)
    ctx.batch_set_wave(args.wave)
    ctx.batch_upload_packed(pts, scans, ins, q, descs)

    # The path's only exchange step: every rank ends up with the per-feature results of all ranks, through the C ABI.
    #   peer (default): k_plane stores its results straight into every rank's gathered block over NVLink (CUDA IPC peer
    #                   memory) and publishes a step counter; no collective is launched
    #   nccl:           one ncclAllGather per step (levins_gather) on the context's stream
    # torch.distributed only carries the 64-byte IPC handles / the 128-byte NCCL id between the ranks.
    nq = n_kf * F
    res_host = {"type": np.zeros(nq, np.int8), "xyz_depth": np.zeros((nq, 4), np.float32)}
    comm = None
    if world > 1:
        if args.gather == "peer":
            mine = torch.from_numpy(ctx.peer_export(world, nq)).to(dev)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            ctx.peer_attach(rank, world, torch.stack(allh).cpu().numpy())
        else:
            uid = torch.from_numpy(abi.Context.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
            dist.broadcast(uid, 0)
            comm = ctx.comm_init(uid.cpu().numpy(), world, rank)
        dist.barrier()

    def gather():
        if comm is not None:
            ctx.gather(comm, world)

    def drain():  # device-side: the stream does not pass this point before every rank's results of the latest step are here
        if world > 1 and comm is None:
            ctx.peer_wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        ctx.batch_run()
        gather()

    def step_e2e():
        ctx.batch_upload_packed(pts, scans, ins, q, descs)
        ctx.batch_run()
        gather()
        drain()
        ctx.batch_download(res_host)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count
        e0.record(stream)
        for _ in range(steps):
            fn()
        drain()  # every step's gather has completed inside the timed region
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launch_count - l0

    # ---- device-resident throughput ----
    for _ in range(max(args.warmup, 3)):
        step_resident()
    drain()
    ctx.batch_download()
    with ClockSampler(local_rank) as clocks:
        ms_total, launches = timed(step_resident, args.steps)
        # keep the sampler alive over the e2e region too so that short runs still collect samples
        e2e_steps = max(3, min(args.steps, args.e2e_steps))
        for _ in range(2):
            step_e2e()
        ms_e2e, _ = timed(step_e2e, e2e_steps)
    ms_per_step = ms_total / args.steps
    total_features = world * nq
    value = total_features * args.steps / (ms_total * 1e-3)
    out = ctx.batch_download()
    n_proj_total = int(out["n_projected"].astype(np.int64).sum())
    pts_per_step = n_proj_total  # per rank
    e2e_value = total_features * e2e_steps / (ms_e2e * 1e-3)
    h2d = int(pts.nbytes + scans.nbytes + ins.nbytes + q.nbytes + descs.nbytes)
    d2h = int(res_host["type"].nbytes + res_host["xyz_depth"].nbytes)

    # ---- per-kernel time, live, with CUDA events on the launching stream ----
    ctx.set_profiling(True)
    prof_ms = np.zeros(4)
    prof_n = np.zeros(4)
    for _ in range(3):
        ctx.batch_run()
        m, n = ctx.batch_kernel_ms()
        prof_ms += m
        prof_n += n
    ctx.set_profiling(False)
    peak, peak_src = load_peaks()
    deskew_launch_ms = prof_ms[1] / max(prof_n[1], 1)
    launches_per_step = prof_n[1] / 3
    alg_bytes_per_launch = 32.0 * n_proj_total / max(launches_per_step, 1)
    achieved = alg_bytes_per_launch / (deskew_launch_ms * 1e-3) / 1e9
    share = prof_ms / max(prof_ms.sum(), 1e-9)
    stage3_ms = float(prof_ms[2] + prof_ms[3]) / 3.0  # k_plane + k_associate per step (3 profiled steps)

    # ---- streaming latency: append scans + fused keyframe through the ABI, wall clock per keyframe ----
    def secondary(fn, *a, **kw):
        """the secondary measurements must never cost the headline line"""
        try:
            return fn(*a, **kw)
        except Exception as ex:  # noqa: BLE001
            return {"error": f"{type(ex).__name__}: {ex}"[:300]}

    p50 = None
    if rank == 0 and not args.no_stream:
        p50 = secondary(stream_latency, abi, synth, local_rank)

    # ---- secondary workload: BASELINE configs[1] as a resident batch (round 1's headline), for continuity ----
    also = None
    if rank == 0 and world == 1 and not args.no_also:
        also = secondary(bench_cfg2, torch, abi, synth, local_rank, stream, peak)

    # ---- BASELINE configs[4]: the offline batch of independent Ouster keyframes, generated on the device from per-keyframe
    #      seeds, sharded over the ranks, with the gather of the depths ----
    cfg5 = None
    if not args.no_cfg5:
        cfg5 = secondary(bench_cfg5, torch, dist, abi, args, rank, world, local_rank, stream, peak)

    # ---- CPU baseline on a bounded sample (rank 0 only), and the parity of the timed batch's results against it ----
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_py as orc

        orc.build()
        cores = host_cores()
        oracle_results = {}
        t, nf, npts, nk = cpu_keyframes(batch, range(n_kf), cores, budget_s=args.cpu_seconds, results=oracle_results)
        parity = parity_report(out, oracle_results, F)
        # the host work levins_append_raw replaces: LidarConverter's loop on one 24 000-point Livox scan (one thread, like
        # the ROS callback)
        from levins import rawfmt

        shp2 = synth.shape(STREAM_CFG)
        kf2 = synth.make_keyframe(STREAM_CFG)
        rec, tb = rawfmt.livox_records(kf2.points[:shp2.pts_per_scan], 2334)
        raw = rawfmt.as_bytes(rec)
        t0c = time.perf_counter()
        for _ in range(20):
            orc.convert_raw(rawfmt.fmt_livox(), raw, len(rec), tb, 0.0, True, 6, 1.0, 200.0)
        conv_us = 1e6 * (time.perf_counter() - t0c) / 20
        cpu = {"value": nf / t, "unit": UNIT, "cores": cores, "kind": "port", "converter_us_per_scan": conv_us,
               "sample": f"{nk} keyframes of the workload ({nf} features, {npts} points) in {t:.1f} s; stage 1 on "
                         f"{cores} threads, stages 2-3 serial + KD-tree like the reference",
               "lidar_pts_per_s": npts / t, "keyframes_per_s": nk / t}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(batch, n_kf),
            "run": {"points_per_keyframe": int(n_proj_total / n_kf), "wave": args.wave,
                    "cells_per_keyframe": int(out["m"].mean()), "associated_fraction": float((out["type"] == 1).mean()),
                    "gather": (args.gather if world > 1 else None),
                    "e2e_input": "packed 16-byte records (x, y, z, u32 ns after the scan stamp), expanded to PointXYZIT on the device"},
            "lidar_pts_per_s": world * pts_per_step * args.steps / (ms_total * 1e-3),
            "keyframes_per_s": world * n_kf * args.steps / (ms_total * 1e-3),
            "stage3_alone": {"features_per_s": world * nq / max(stage3_ms * 1e-3, 1e-12), "ms_per_step": stage3_ms,
                             "kernels": "k_associate + k_plane, live CUDA events (SURVEY 8d: features/s also for stage 3 alone; in the "
                                        "fused path k_associate deskews the winners of the cells it visits)"},
            "p50_keyframe_ms": p50.get("pageable") if p50 else None,
            "p50_keyframe_ms_pinned_scans": p50.get("pinned") if p50 else None,
            "streaming": p50,
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_deskew_bin", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": load_traffic(),
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch, "launch_ms": deskew_launch_ms,
                         "peak_source": peak_src,
                         "kernel_time_share": {"k_prepare": share[0], "k_deskew_bin": share[1], "k_plane": share[2],
                                               "k_associate": share[3]},
                         "whole_step_frac": (32.0 * n_proj_total / (ms_per_step * 1e-3) / 1e9) / peak},
            "cpu_baseline": cpu,
            "parity": parity,
            "also": also,
            "cfg5_offline_batch": cfg5,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if comm is not None:
            abi.Context.comm_destroy(comm)
        else:
            ctx.peer_detach()
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def bench_cfg5(torch, dist, abi, args, rank, world, device, stream, peak, steps=5):
    """BASELINE configs[4]: an offline batch of independent Ouster OS1-128 keyframes (the reference's default 0.5 s stack
    window: 5 scans = 655 360 points, 300 features each), 10 000 of them at 8 GPUs = 1 250 per GPU (weak scaling at other N).
    The keyframes are GENERATED ON THE DEVICE from per-keyframe seeds (le-vins_b200/harness/synthgen.cu) and attached without a
    copy; the per-feature results of all ranks are exchanged through the C ABI.  Rank 0 re-generates a few keyframes on the
    host with the generator's numpy twin (bit-identical, tests/test_synthgen.py) and checks the GPU results against the oracle."""
    from levins import synthgen

    sp = synthgen.Spec()
    n_kf = args.cfg5_keyframes
    first = rank * n_kf
    gen = synthgen.DeviceGenerator(torch, device, sp)
    d_pts, d_ins, d_q = gen.generate(first, n_kf)
    ctx = abi.Context(device, stack_interval=sp.stack_interval)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_extrinsics(sp.rig.pose_b_c, sp.rig.pose_c_l, sp.td_b_l)
    ctx.batch_attach_device(d_pts.data_ptr(), n_kf * sp.per_kf, d_ins.data_ptr(), n_kf * sp.n_ins, d_q.data_ptr(), n_kf * sp.n_feat,
                            gen.descs(first, n_kf))
    nq = n_kf * sp.n_feat
    comm = None
    if world > 1:
        if args.gather == "peer":
            mine = torch.from_numpy(ctx.peer_export(world, nq)).to(d_q.device)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            ctx.peer_attach(rank, world, torch.stack(allh).cpu().numpy())
        else:
            uid = torch.from_numpy(abi.Context.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(d_q.device)
            dist.broadcast(uid, 0)
            comm = ctx.comm_init(uid.cpu().numpy(), world, rank)
        dist.barrier()

    def step():
        ctx.batch_run()
        if comm is not None:
            ctx.gather(comm, world)

    def finish():
        if world > 1 and comm is None:
            ctx.peer_wait()

    for _ in range(2):
        step()
    finish()
    ctx.batch_download()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    finish()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=d_q.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    out = ctx.batch_download()
    gathered_ok = None
    if world > 1:  # what arrived from the other ranks is what they computed: compare rank 0's slot everywhere via a checksum
        g = ctx.gathered_download(world)
        mine_ok = bool(np.array_equal(g["type"][rank, :nq], out["type"]) and g["xyz_depth"][rank, :nq].tobytes() == out["xyz_depth"].tobytes())
        sums = torch.tensor([float(g["xyz_depth"][r, :nq, 3].astype(np.float64).sum()) for r in range(world)], dtype=torch.float64,
                            device=d_q.device)
        ref = sums.clone()
        dist.broadcast(ref, 0)
        ok = torch.tensor([1.0 if (mine_ok and bool(torch.equal(sums, ref))) else 0.0], dtype=torch.float64, device=d_q.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        gathered_ok = bool(ok.item() == 1.0)
    parity = None
    if rank == 0 and not args.no_cpu:
        import oracle_py as orc

        orc.build()
        dirs = synthgen.directions()
        res = {}
        for k in (0, n_kf // 2, n_kf - 1):
            pts, ins, q = synthgen.keyframe_numpy(sp, first + k, dirs)
            res[k] = orc.keyframe(pts, sp.stamp(first + k), ins, sp.rig.pose_b_c, sp.rig.pose_c_l, sp.td_b_l, q, sp.stack_interval,
                                  nthreads=host_cores())
        parity = parity_report(out, res, sp.n_feat)
    npts = int(out["n_projected"].astype(np.int64).sum())
    if world > 1:
        dist.barrier()
        if comm is not None:
            abi.Context.comm_destroy(comm)
        else:
            ctx.peer_detach()
        dist.barrier()
    ctx.close()
    del d_pts, d_ins, d_q
    return {"workload": f"cfg5_offline_os1_128_x{sp.scans_per_kf}scans_F{sp.n_feat}_{n_kf}_keyframes_per_gpu_generated_on_device",
            "keyframes_total": world * n_kf, "ms_per_step": ms, "keyframes_per_s": world * n_kf / (ms * 1e-3),
            "features_per_s": world * nq / (ms * 1e-3), "lidar_pts_per_s": world * npts / (ms * 1e-3),
            "points_per_keyframe": npts // n_kf, "cells_per_keyframe": int(out["m"].mean()),
            "associated_fraction": float((out["type"] == 1).mean()),
            "whole_step_hbm_frac": 32.0 * npts / (ms * 1e-3) / 1e9 / peak, "resident_point_bytes_per_gpu": int(d_bytes(n_kf, sp)),
            "gather": (args.gather if world > 1 else None), "gathered_blocks_consistent": gathered_ok, "parity": parity}


def d_bytes(n_kf, sp):
    return n_kf * sp.per_kf * 32


def bench_cfg2(torch, abi, synth, device, stream, peak, n_kf=256, steps=10):
    """Device-resident throughput of round 1's headline workload (R3LIVE handheld shape, 256 keyframes x 120 001 points x
    300 features: every point lies inside the camera's field of view)."""
    b = cached_batch(synth, f"cfg2_k{n_kf}_s{synth.BASE_SEED + 5}_u48",
                     lambda: synth.make_batch(2, n_kf, seed=synth.BASE_SEED + 5, n_unique_scans=48))
    ctx = abi.Context(device)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_extrinsics(b.shape.rig.pose_b_c, b.shape.rig.pose_c_l, b.shape.rig.td_b_l)
    ctx.batch_upload(b.points, b.ins, b.queries.reshape(-1, 3), abi.batch_descs(b))
    for _ in range(3):
        ctx.batch_run()
        ctx.batch_download()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ctx.batch_run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = ctx.batch_download()
    npts = int(out["n_projected"].astype(np.int64).sum())
    ctx.set_profiling(True)
    ctx.batch_run()
    kms, kn = ctx.batch_kernel_ms()
    ctx.close()
    return {"workload": f"cfg2_r3live_avia_24kpts_x5scans_1280x1024_F{b.queries.shape[1]}_batch{n_kf}_independent_keyframes",
            "ms_per_step": ms, "lidar_pts_per_s": npts / (ms * 1e-3), "keyframes_per_s": n_kf / (ms * 1e-3),
            "features_per_s": n_kf * b.queries.shape[1] / (ms * 1e-3), "points_per_keyframe": npts // n_kf,
            "cells_per_keyframe": int(out["m"].mean()), "associated_fraction": float((out["type"] == 1).mean()),
            "whole_step_hbm_frac": 32.0 * npts / (ms * 1e-3) / 1e9 / peak,
            "k_deskew_bin_hbm_frac": 32.0 * npts / (float(kms[1]) * 1e-3) / 1e9 / peak,
            "kernel_ms": {"k_prepare": float(kms[0]), "k_deskew_bin": float(kms[1]), "k_plane": float(kms[2]),
                          "k_associate": float(kms[3])}}


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f).get("k_deskew_bin_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def stream_latency(abi, synth, device, n_keyframes=40):
    """Config-2 streaming: one scan appended per 0.1 s of sensor time, one keyframe every 0.5 s.  Wall time from the
    first append of a keyframe's new scans to its results on the host; p50 over the keyframes.  Four ways of handing the
    scans over: PointXYZIT in ordinary pageable memory (what the unmodified ROS converter produces), PointXYZIT in
    page-locked buffers from levins_alloc_host, and the driver's own Livox records (20 B / point, converted on the device
    by levins_append_raw) from pageable and from page-locked memory."""
    from levins import rawfmt

    shp = synth.shape(STREAM_CFG)
    b = synth.make_batch(STREAM_CFG, n_keyframes, n_unique_scans=12, seed=synth.BASE_SEED + 77)
    per = shp.pts_per_scan
    n_scans = len(b.points) // per
    week = 2334
    recs, bases = [], []
    for s in range(n_scans):
        r, tb = rawfmt.livox_records(b.points[s * per:(s + 1) * per], week)
        recs.append(rawfmt.as_bytes(r))
        bases.append(tb)
    raw_all = np.concatenate(recs)
    step = per * rawfmt.LIVOX.itemsize
    out = {}
    for kind in ("pageable", "pinned", "livox_pageable", "livox_pinned"):
        raw = kind.startswith("livox")
        src, hb = (raw_all if raw else b.points), None
        if kind.endswith("pinned"):
            hb = abi.HostBuffer(len(src), src.dtype)
            hb.array[:] = src
            src = hb.array
        ctx = abi.Context(device)
        ctx.set_extrinsics(shp.rig.pose_b_c, shp.rig.pose_c_l, shp.rig.td_b_l)
        times, t_append, t_call = [], [], []
        scan = 0
        for k, stamp in enumerate(b.stamps):
            kf_ins = b.ins[b.ins_begin[k]: b.ins_begin[k] + b.ins_count[k]]
            t0 = time.perf_counter()
            while scan < n_scans and b.points["time"][scan * per] <= stamp + 1e-9:  # everything that started before the stamp
                ta = time.perf_counter()
                if raw:
                    ctx.append_raw(rawfmt.fmt_livox(), src[scan * step:(scan + 1) * step], per, timebase_ns=bases[scan], scan_line=6,
                                   want_result=False)
                else:
                    ctx.append_scan(src[scan * per:(scan + 1) * per])
                t_append.append(1e6 * (time.perf_counter() - ta))
                scan += 1
            if kind == "pinned":  # live operation: the scans arrived earlier, only the keyframe call itself is on the clock
                ctx.synchronize()
                tk = time.perf_counter()
                ctx.keyframe(float(stamp), kf_ins, b.queries[k])
                t_call.append(1e3 * (time.perf_counter() - tk))
                t0 += 0.0
            else:
                ctx.keyframe(float(stamp), kf_ins, b.queries[k])
            times.append(1e3 * (time.perf_counter() - t0))
        ctx.close()
        if hb is not None:
            src = None
            hb.close()
        out[kind] = float(np.median(times[3:])) if len(times) > 3 else None
        out[kind + "_append_us"] = float(np.median(t_append[5:])) if len(t_append) > 5 else None
        if t_call:
            out["keyframe_call_ms_scans_resident"] = float(np.median(t_call[3:]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--keyframes", type=int, default=DEFAULT_KEYFRAMES, help="keyframes per GPU per step")
    ap.add_argument("--wave", type=int, default=0, help="keyframes per wave of launches (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the device-generated offline batch (configs[4])")
    ap.add_argument("--cfg5-keyframes", type=int, default=1250, help="keyframes per GPU of the offline batch (10 000 at 8 GPUs)")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary measurement (configs[1] as a resident batch)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="exchange step at N > 1 (both through the C ABI)")
    ap.add_argument("--no-stream", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
