// kernels.cuh -- the kernels of the LiDAR depth-enhancement path (sm_100a, hand-written, no libraries).
//
//   k_prepare      grid (keyframes, slices): window search (le_vins.cc:413-437), keyframe pose (le_vins.cc:441-442), the
//                  per-INS-interval table that replaces the per-point pose interpolation (misc.cc:62-103) with its
//                  frustum-cull records, and the chunk descriptors of k_deskew_bin
//   k_deskew_bin   per point: range gate + deskew (le_vins.cc:445-464) fused with range-image binning
//                  (tracking.cc:307-336) through a 64-bit atomicMin key (range bits | point index); persistent warps,
//                  each with its own TMA ring; certified frustum cull in front of the f64 algebra
//   k_associate    per feature (one warp): [undistort + pixel2unitcam (camera.cc:72-74,123-135)], exact windowed 5-NN
//                  (tracking.cc:571); in the fused path it also turns the z-buffer winners of its window into unit
//                  vectors + ranges (tracking.cc:340-354) on the fly
//   k_plane        per feature (one thread): threshold, QR plane fit, validity checks, depth (tracking.cc:573-631),
//                  [accept gate + cam2world (tracking.cc:431-439)], [stores into every rank's gathered block: exchange.cuh]
//   k_count_cells  M per keyframe (side stream)
//   k_finalize     stage-wise API only: per cell winner -> unit vector + range
//   k_raw_count / k_raw_scatter   driver records -> PointXYZIT in the ring + ring bookkeeping (lidar_converter.cc:35-256)
//   k_unpack       packed 16-byte host records -> PointXYZIT (batch upload)
//   k_plane_estimation / k_point_to_plane_std / k_project   PointCloudCommon (pointcloud.cc:50-190), batched
//   k_depth_factor / k_chi2   depth factor + chi-square culling of the optimiser (depth_factor.h:38-42, optimizer.cc:752-783)
// plus k_bin_cloud / k_compact / k_ring_append for the stage-wise ABI entry points and k_peer_signal / k_peer_wait (exchange.cuh).
#pragma once

#include <type_traits>

#include "device_math.cuh"
#include "exchange.cuh"

namespace lv {

#ifndef LV_THREADS
#define LV_THREADS 512 // one 16-warp CTA per SM: 2 % faster than two 8-warp CTAs (profiles/r01_experiments.md)
#endif
constexpr int DESKEW_THREADS = LV_THREADS;
constexpr int FINALIZE_THREADS = 256;
#ifndef LV_ASSOC_WARPS
#define LV_ASSOC_WARPS 2
#endif
#ifndef LV_ASSOC_MINBLK
#define LV_ASSOC_MINBLK 12
#endif
constexpr int ASSOC_WARPS = LV_ASSOC_WARPS;
constexpr int WIN_ROWS = 4; // |d row| <= 4, |d col| <= 6: every point within the acceptance radius 3*sin(0.3 deg)
constexpr int WIN_COLS = 6; // of a query lies in this 9 x 13 window ("windowed k-NN is exact")
constexpr int WIN_W = 2 * WIN_COLS + 1;
constexpr int WIN_CELLS = (2 * WIN_ROWS + 1) * WIN_W; // 117

struct RingState {
    unsigned long long head;  // physical index of the front point
    unsigned long long count; // live points
};

__device__ __forceinline__ unsigned long long wrap(unsigned long long i, unsigned long long cap) { return i >= cap ? i - cap : i; }

// Smallest i in [0, count] with pred(i) true, for a predicate that is monotone (false ... false true ... true).
// 32-ary search: one probe per lane and round, so a 1 M-point window needs 4 dependent global loads, not 20.
template <typename Pred>
__device__ __forceinline__ unsigned long long warp_search(unsigned long long count, Pred pred) {
    const unsigned lane = threadIdx.x & 31;
    unsigned long long lo = 0, hi = count;
    while (hi > lo) {
        const unsigned long long span = hi - lo;
        const unsigned long long step = (span + 31) / 32;
        unsigned long long off = (lane + 1) * step;
        if (off > span) off = span;
        const unsigned long long probe = lo + off - 1;
        const bool p = pred(probe);
        const unsigned ballot = __ballot_sync(0xffffffffu, p);
        if (ballot == 0) {
            lo = hi;
        } else {
            const int fl = __ffs(ballot) - 1;
            const unsigned long long new_hi = __shfl_sync(0xffffffffu, probe, fl);
            const unsigned long long prev = __shfl_sync(0xffffffffu, probe, fl > 0 ? fl - 1 : 0);
            if (fl > 0) lo = prev + 1;
            hi = new_hi;
        }
    }
    return hi;
}

// ------------------------------------------------------------------------------------------------
// ring buffer append bookkeeping: VINS::addNewPointCloud eviction rule (le_vins.cc:156-159)
// ------------------------------------------------------------------------------------------------
// (one warp)
__device__ __forceinline__ void ring_append(RingState* ring, const PointXYZIT* __restrict__ pts, unsigned long long cap,
                                            unsigned long long n_new, double reserved_time);
__global__ void k_ring_append(RingState* ring, const PointXYZIT* __restrict__ pts, unsigned long long cap,
                              unsigned long long n_new, double reserved_time) {
    ring_append(ring, pts, cap, n_new, reserved_time);
}
__device__ __forceinline__ void ring_append(RingState* ring, const PointXYZIT* __restrict__ pts, unsigned long long cap,
                                            unsigned long long n_new, double reserved_time) {
    const unsigned long long head = ring->head;
    const unsigned long long count = ring->count + n_new;
    if (count == 0) return;
    const double t_back = pts[wrap(head + count - 1, cap)].time;
    // pop_front while |back - front| > reserved  <=>  first i with |back - t_i| <= reserved (time-sorted buffer)
    const unsigned long long pop = warp_search(count, [&](unsigned long long i) {
        return !(fabs(t_back - pts[wrap(head + i, cap)].time) > reserved_time);
    });
    if ((threadIdx.x & 31) == 0) {
        ring->head = wrap(head + pop, cap);
        ring->count = count - pop;
    }
}

// ------------------------------------------------------------------------------------------------
// Driver records -> PointXYZIT straight into the ring (ROS/lidar/lidar_converter.cc:35-256,
// le_vins/common/gpstime.h:36-41).  Three small kernels: per-block survivor counts + start/end, an exclusive scan over
// the blocks, and the ordered scatter.  point_filter_num_ is the constant 1 in the reference (lidar_converter.h:61).
// ------------------------------------------------------------------------------------------------
struct RawFormat {
    unsigned kind, point_step, off_x, off_y, off_z, off_intensity, off_time, off_ring, off_tag, time_kind, ring_bytes, pad;
};
struct RawResult {
    unsigned long long count;
    double start, end;
};
struct RawArgs {
    const unsigned char* rec;
    unsigned n;
    RawFormat f;
    unsigned long long timebase_ns;
    double stamp; // PointCloud2 kinds with a header-relative time: already converted to GPS time by the host
    int to_gps, scan_line;
    double near2, far2;
    unsigned* block_kept;       // [ceil(n / 256)] survivor counts, turned into exclusive offsets by k_raw_scan
    unsigned long long* minmax; // ordered bit patterns of start (min) and end (max)
    PointXYZIT* ring;
    unsigned long long cap;
    RawResult* result;
    RingState* ring_state; // the scatter appends behind head + count and the last block to finish does the bookkeeping
    unsigned* ticket;      // blocks finished (reset by k_raw_count)
    double reserved_time;  // le_vins.cc:156
};
constexpr int RAW_BLOCK = 256;

__device__ __forceinline__ unsigned raw_u32(const unsigned char* p) { // little-endian, any alignment
    return (unsigned) p[0] | ((unsigned) p[1] << 8) | ((unsigned) p[2] << 16) | ((unsigned) p[3] << 24);
}
__device__ __forceinline__ double unix2gps_sow(double unixs) { // gpstime.h:36-41, GPS_LEAP_SECOND = 18
    const double seconds = unixs + 18 - 315964800;
    const int week = (int) floor(seconds / 604800);
    return seconds - week * 604800;
}
// doubles -> unsigned keys with the same order (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned long long ordered_key(double d) {
    const unsigned long long b = (unsigned long long) __double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_value(unsigned long long k) {
    return __longlong_as_double((long long) ((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
// one record: false if the converter skips it before looking at the range (Livox tag / line); `kept` = survives the range gate
__device__ __forceinline__ bool raw_decode(const RawArgs& a, unsigned k, PointXYZIT& p, bool& kept) {
    const unsigned char* r = a.rec + (size_t) k * a.f.point_step;
    kept = false;
    p.pad0 = 0.f;
    p.pad1 = 0.f;
    if (a.f.kind == 1) {
        const unsigned line = r[a.f.off_ring], tag = r[a.f.off_tag];
        if (!((int) line < a.scan_line && ((tag & 0x30u) == 0x00u || (tag & 0x30u) == 0x10u))) return false; // :55
        p.intensity = (float) r[a.f.off_intensity];                                                          // :244
        p.time = (double) (a.timebase_ns + (unsigned long long) raw_u32(r + a.f.off_time)) * 1.0e-9;         // :245
        if (a.to_gps) p.time = unix2gps_sow(p.time);
    } else {
        p.intensity = __uint_as_float(raw_u32(r + a.f.off_intensity));
        if (a.f.time_kind == 1) {
            p.time = a.stamp + (double) __uint_as_float(raw_u32(r + a.f.off_time)); // :113
        } else if (a.f.time_kind == 2) {
            p.time = a.stamp + (double) raw_u32(r + a.f.off_time) * 1.0e-9; // :162
        } else {
            const unsigned long long b = (unsigned long long) raw_u32(r + a.f.off_time) | ((unsigned long long) raw_u32(r + a.f.off_time + 4) << 32);
            p.time = __longlong_as_double((long long) b); // :213-221
            if (a.to_gps) p.time = unix2gps_sow(p.time);
        }
    }
    p.x = __uint_as_float(raw_u32(r + a.f.off_x));
    p.y = __uint_as_float(raw_u32(r + a.f.off_y));
    p.z = __uint_as_float(raw_u32(r + a.f.off_z));
    const double sq = (double) v3f_sqnorm(p.x, p.y, p.z); // :66
    kept = (sq > a.near2) && (sq < a.far2);               // :67
    return true;
}

__global__ void __launch_bounds__(RAW_BLOCK) k_raw_count(const __grid_constant__ RawArgs a) {
    const unsigned k = blockIdx.x * RAW_BLOCK + threadIdx.x;
    PointXYZIT p;
    bool kept = false, seen = false;
    if (k < a.n) seen = raw_decode(a, k, p, kept);
    const int c = __syncthreads_count(kept);
    if (threadIdx.x == 0) a.block_kept[blockIdx.x] = (unsigned) c;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.ticket = 0;
    // start / end over every considered record, before the range gate (:58-63)
    unsigned long long lo = seen ? ordered_key(p.time) : ~0ull, hi = seen ? ordered_key(p.time) : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo != ~0ull) atomicMin(a.minmax, lo);
        if (hi != 0ull) atomicMax(a.minmax + 1, hi);
    }
}

// Ordered scatter.  Every block sums the survivor counts of the blocks before it (a few hundred values from L2), places its
// survivors behind the ring's tail, and the last block to finish writes the result record and does VINS::addNewPointCloud's
// bookkeeping (count += appended, 11 s eviction): conversion + append are two launches in all.
__global__ void __launch_bounds__(RAW_BLOCK) k_raw_scatter(const __grid_constant__ RawArgs a) {
    __shared__ unsigned s_warp[RAW_BLOCK / 32];
    __shared__ unsigned s_base[RAW_BLOCK / 32];
    __shared__ unsigned s_last;
    const unsigned k = blockIdx.x * RAW_BLOCK + threadIdx.x;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PointXYZIT p;
    bool kept = false;
    if (k < a.n) raw_decode(a, k, p, kept);
    // exclusive offset of this block
    unsigned part = 0;
    for (unsigned i = threadIdx.x; i < blockIdx.x; i += RAW_BLOCK) part += a.block_kept[i];
    part = __reduce_add_sync(0xffffffffu, part);
    const unsigned bal = __ballot_sync(0xffffffffu, kept);
    if (lane == 0) {
        s_warp[warp] = __popc(bal);
        s_base[warp] = part;
    }
    __syncthreads();
    unsigned off = 0;
#pragma unroll
    for (int w = 0; w < RAW_BLOCK / 32; w++) off += s_base[w];
    for (unsigned w = 0; w < warp; w++) off += s_warp[w];
    const unsigned long long tail = wrap(a.ring_state->head + a.ring_state->count, a.cap); // stable until the last block updates it
    if (kept) {
        const unsigned long long dst = wrap(tail + off + __popc(bal & ((1u << lane) - 1u)), a.cap);
        float4* out = reinterpret_cast<float4*>(a.ring + dst);
        out[0] = make_float4(p.x, p.y, p.z, 0.f);
        reinterpret_cast<float2*>(out + 1)[0] = make_float2(p.intensity, 0.f);
        reinterpret_cast<double*>(out + 1)[1] = p.time;
    }
    // last block: result record + ring bookkeeping
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last || warp != 0) return;
    __threadfence();
    unsigned total = 0;
    for (unsigned i = lane; i < gridDim.x; i += 32) total += __ldcg(a.block_kept + i);
    total = __reduce_add_sync(0xffffffffu, total);
    if (lane == 0) {
        a.result->count = total;
        a.result->start = ordered_value(__ldcg(a.minmax));
        a.result->end = ordered_value(__ldcg(a.minmax + 1));
        // start = DBL_MAX, end = 0 for the next scan (lidar_converter.cc:50-51)
        a.minmax[0] = ordered_key(DBL_MAX);
        a.minmax[1] = ordered_key(0.0);
    }
    __syncwarp();
    ring_append(a.ring_state, a.ring, a.cap, total, a.reserved_time);
}

// ------------------------------------------------------------------------------------------------
// k_unpack: packed 16-byte points (x, y, z, u32 nanoseconds after the scan's stamp: the fields of a driver record this path
// reads, with the time as the Ouster / Livox drivers deliver it) -> PointXYZIT in the batch pool.  time = stamp + dt * 1e-9
// (ROS/lidar/lidar_converter.cc:162, :245).  Half the PCIe bytes of uploading PointXYZIT records.
// ------------------------------------------------------------------------------------------------
struct PackedPoint {
    float x, y, z;
    unsigned dt_ns;
};
struct ScanDesc {
    unsigned long long first, count;
    double stamp;
};
__global__ void __launch_bounds__(256) k_unpack(const PackedPoint* __restrict__ in, unsigned long long n, const ScanDesc* __restrict__ scans,
                                                unsigned n_scans, PointXYZIT* __restrict__ out) {
    const unsigned long long k = (unsigned long long) blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    // the scan that holds point k: last scan with first <= k (scans are sorted by `first`; a block rarely spans two)
    unsigned lo = 0, hi = n_scans;
    while (hi - lo > 1) {
        const unsigned mid = (lo + hi) >> 1;
        if (__ldg(&scans[mid].first) <= k) lo = mid; else hi = mid;
    }
    const double stamp = __ldg(&scans[lo].stamp);
    const uint4 p = __ldg(reinterpret_cast<const uint4*>(in + k));
    float4* o = reinterpret_cast<float4*>(out + k);
    o[0] = make_float4(__uint_as_float(p.x), __uint_as_float(p.y), __uint_as_float(p.z), 0.f);
    const double t = stamp + (double) p.w * 1.0e-9;
    o[1] = make_float4(0.f, 0.f, __int_as_float(__double2loint(t)), __int_as_float(__double2hiint(t)));
}

// ------------------------------------------------------------------------------------------------
// chunks: k_deskew_bin's unit of work is 128 consecutive points of one keyframe (one warp, one iteration)
// ------------------------------------------------------------------------------------------------
constexpr int DESKEW_PPT = 4;              // points per thread and chunk
constexpr int CHUNK_PTS = 32 * DESKEW_PPT; // 128 points = 4 KB
// what lane 0 of the owning warp needs to fetch one chunk
struct __align__(16) ChunkDesc {
    unsigned long long phys; // physical index of the chunk's first point in the pool / ring
    unsigned ent0;           // absolute index (into the interval-table pool) of the entry of the chunk's first point
    unsigned cnt;            // points in the chunk (0 = empty)
};

// Index of the INS interval that holds time t (misc.cc:25-60 semantics on the IMU times of the window): i in [1, n - 1] with
// time[i-1] <= t < time[i], 0 outside the window or when no pair brackets t.  Near-uniform sampling: the guess from the mean
// spacing is almost always right; otherwise bisection.
__device__ __forceinline__ int interval_of_time(const InsSample* __restrict__ w, int nins, double t, double t0, double t_last, double inv_mean_dt) {
    if (nins < 2 || t0 > t || t_last <= t) return 0; // misc.cc:28-30
    int g = 1 + (int) ((t - t0) * inv_mean_dt);
    g = min(max(g, 1), nins - 1);
    if (w[g - 1].time <= t && t < w[g].time) return g;
    int lo = 1, hi = nins - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (w[mid].time > t) hi = mid; else lo = mid + 1;
    }
    return (w[lo - 1].time <= t && t < w[lo].time) ? lo : 0;
}

// Descriptor of chunk `c` of a keyframe.  The entry index is only a hint for k_deskew_bin (every point still checks
// that its entry brackets it), so any answer is correct; the right one is fast.  k_deskew_bin stages entries idx and
// idx + 1: the next interval, the fallback copy in slot nins after the last one, the first interval for a chunk that
// starts before the window.
__device__ __forceinline__ ChunkDesc chunk_desc_from_time(unsigned c, unsigned n_proj, unsigned long long phys, double t_point, double td_b_l,
                                                          const InsSample* __restrict__ w, unsigned long long table_begin, int nins,
                                                          double t0, double t_last, double inv_mean_dt) {
    ChunkDesc d;
    d.phys = 0;
    d.ent0 = 0;
    d.cnt = 0;
    const unsigned off = c * CHUNK_PTS;
    if (off < n_proj) {
        const int idx = interval_of_time(w, nins, t_point + td_b_l /* le_vins.cc:455 */, t0, t_last, inv_mean_dt);
        d.phys = phys;
        d.ent0 = (unsigned) (table_begin + idx);
        d.cnt = min((unsigned) CHUNK_PTS, n_proj - off);
    }
    return d;
}

// ------------------------------------------------------------------------------------------------
// k_prepare: grid (keyframes, slices).  Every block of a keyframe repeats the three window searches (a handful of dependent
// loads) and the keyframe pose; the INS-interval table and the chunk descriptors are split over the slices.
// ------------------------------------------------------------------------------------------------
struct PrepareArgs {
    KfState* kfs;
    const PointXYZIT* pts;
    unsigned long long pts_cap; // ring capacity (ULLONG_MAX for linear pools)
    const InsSample* ins;
    Interval* table;
    ChunkDesc* chunks; // [gridDim.x][chunks_per_kf]
    unsigned chunks_per_kf;
    RingState* ring; // streaming: the keyframe's buffer is the ring (le_vins.cc:410-437), and is popped
    Extrinsics ext;
    Consts c;
};

// The certified-cull record of an interval entry (see Interval): the deskew frozen at the middle of the entry's own range
// of s = (t - ts0) * inv_dts, and a bound on everything that was dropped.  With Rod(k, phi) = I + sin(phi) K + (1 - cos(phi)) K^2:
//   p_c0(s) = A Rod(k, -s theta) (R_bc (R_cl p + t_cl) + t_bc) + c0 + s c1
//   |Rod(k, a) - Rod(k, b)| <= |a - b| (|k| + |k|^2)                      (operator norm; |k| = 1 for a proper axis)
// so |p_c0(s) - p_c0(s_mid)| <= h (|A| theta kk (|R_bc| (|R_cl| |p| + |t_cl|) + |t_bc|) + |c1|),  h = max |s - s_mid|,
// with Frobenius norms standing in for operator norms.  On top of that come the f32 evaluation of M p + m (4 roundings per
// row, M and m themselves rounded), the intermediate f32 rounding of pointcloud.cc:68 and the f64 round-off of the real path:
// all far below the 1e-5 (|M| |p| + |m|) that is added for them.
__device__ __forceinline__ void fill_cull_record(Interval& e, const Extrinsics& ex, bool fallback) {
    double s_mid = 0.0, h = 0.0;
    if (!fallback) {
        const double s_lo = (e.t_lo - e.ts0) * e.inv_dts, s_hi = (e.t_hi - e.ts0) * e.inv_dts;
        s_mid = 0.5 * (s_lo + s_hi);
        h = 0.5 * fabs(s_hi - s_lo) * 1.001 + 1e-9;
    }
    double sn, omc;
    sin_omc(-s_mid * e.theta, sn, omc);
    const double kx = e.k[0], ky = e.k[1], kz = e.k[2];
    // Rod = I + sn K + omc K^2, row-major
    const double K[9] = {0, -kz, ky, kz, 0, -kx, -ky, kx, 0};
    double Rod[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double k2 = (K[3 * r] * K[c] + K[3 * r + 1] * K[3 + c]) + K[3 * r + 2] * K[6 + c];
            Rod[3 * r + c] = (r == c ? 1.0 : 0.0) + sn * K[3 * r + c] + omc * k2;
        }
    double AR[9], ARb[9], Md[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) AR[3 * r + c] = (e.A[3 * r] * Rod[c] + e.A[3 * r + 1] * Rod[3 + c]) + e.A[3 * r + 2] * Rod[6 + c];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) // R_bc is column-major: element (i, c) = Rbc[3 * c + i]
            ARb[3 * r + c] = (AR[3 * r] * ex.Rbc[3 * c] + AR[3 * r + 1] * ex.Rbc[3 * c + 1]) + AR[3 * r + 2] * ex.Rbc[3 * c + 2];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            Md[3 * r + c] = (ARb[3 * r] * ex.Rcl[3 * c] + ARb[3 * r + 1] * ex.Rcl[3 * c + 1]) + ARb[3 * r + 2] * ex.Rcl[3 * c + 2];
    double md[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
        md[r] = ((ARb[3 * r] * ex.tcl[0] + ARb[3 * r + 1] * ex.tcl[1]) + ARb[3 * r + 2] * ex.tcl[2]) +
                ((AR[3 * r] * ex.tbc[0] + AR[3 * r + 1] * ex.tbc[1]) + AR[3 * r + 2] * ex.tbc[2]) + (e.c0[r] + s_mid * e.c1[r]);
    auto fro = [](const double* m, int n) {
        double s = 0;
        for (int i = 0; i < n; i++) s += m[i] * m[i];
        return sqrt(s);
    };
    const double nA = fro(e.A, 9), nRbc = fro(ex.Rbc, 9), nRcl = fro(ex.Rcl, 9), nk = fro(e.k, 3);
    const double rot = nA * h * fabs(e.theta) * (nk + nk * nk);
    double ca = rot * nRbc * nRcl + 1e-5 * fro(Md, 9);
    double cb = rot * (nRbc * fro(ex.tcl, 3) + fro(ex.tbc, 3)) + h * fro(e.c1, 3) + 1e-5 * fro(md, 3) + 1e-6;
    ca *= 1.01;
    cb *= 1.01;
    bool ok = (ca == ca) && (cb == cb) && ca < 0.05 && cb < 1e6;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        e.M[i] = (float) Md[i];
        ok = ok && (fabs(Md[i]) < 1e6);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        e.m[i] = (float) md[i];
        ok = ok && (fabs(md[i]) < 1e9);
    }
    e.ca = ok ? (float) ca * 1.000001f : -1.0f; // (float) rounds to nearest: push it up again
    e.cb = ok ? (float) cb * 1.000001f : 0.0f;
    e.pad[0] = e.pad[1] = 0.f;
}

#ifndef LV_PREP_MINBLK
#define LV_PREP_MINBLK 2
#endif
#ifndef LV_PREP_CHUNKS // chunk descriptors per k_prepare block (decides the number of slices per keyframe)
#define LV_PREP_CHUNKS 4096
#endif
__global__ void __launch_bounds__(256, LV_PREP_MINBLK) k_prepare(const __grid_constant__ PrepareArgs a) {
    KfState* kf = a.kfs + blockIdx.x;
    __shared__ double sR0[9], st0[3];
    __shared__ unsigned long long s_first, s_end, s_begin, s_count;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned slice = blockIdx.y, n_slices = gridDim.y;

    if (threadIdx.x == 0) {
        if (a.ring) { // (the ring state is only changed by the last block of the grid to finish, see below)
            s_begin = a.ring->head;
            s_count = a.ring->count;
        } else {
            s_begin = kf->pt_begin;
            s_count = kf->pt_count;
        }
    }
    __syncthreads();
    const unsigned long long begin = s_begin, count = s_count;
    const InsSample* __restrict__ w = a.ins + kf->ins_begin;
    const int nins = (int) kf->ins_count;
    const double stamp = kf->stamp;
    const double last_stamp = stamp - a.c.stack_interval; // le_vins.cc:395

    if (warp == 0) { // le_vins.cc:413-424: pop while front.time < last_stamp
        const unsigned long long first =
            warp_search(count, [&](unsigned long long i) { return !(a.pts[wrap(begin + i, a.pts_cap)].time < last_stamp); });
        if (lane == 0) s_first = first;
    } else if (warp == 1) { // le_vins.cc:431-437: first k with time > curr_stamp
        const unsigned long long end =
            warp_search(count, [&](unsigned long long i) { return a.pts[wrap(begin + i, a.pts_cap)].time > stamp; });
        if (lane == 0) s_end = end;
    } else if (warp == 2) { // le_vins.cc:441-442: last_keyframe_pose = getPoseFromInsWindow(curr_stamp)
        unsigned long long idx = 0;
        if (!(nins == 0 || w[0].time > stamp || w[nins - 1].time <= stamp))
            idx = warp_search((unsigned long long) nins, [&](unsigned long long i) { return w[i].time > stamp; });
        if (lane == 0) {
            double R0[9], t0[3];
            if (idx > 0)
                interpolate_pose(w[idx - 1], w[idx], stamp, a.ext.Rbc, a.ext.tbc, R0, t0);
            else
                state_pose(w[nins - 1], a.ext.Rbc, a.ext.tbc, R0, t0);
#pragma unroll
            for (int i = 0; i < 9; i++) sR0[i] = R0[i];
            st0[0] = t0[0];
            st0[1] = t0[1];
            st0[2] = t0[2];
        }
    }
    __syncthreads();

    const unsigned long long first = s_first;
    const unsigned long long end = s_end > first ? s_end : first;
    const unsigned n_proj = (unsigned) (end - first);
    const unsigned long long proj_begin = wrap(begin + first, a.pts_cap); // the projected slice starts at the first kept point
    const double t_first = w[0].time, t_last = w[nins - 1].time;
    const double inv_mean_dt = nins > 1 ? (double) (nins - 1) / (t_last - t_first) : 0.0;
    if (slice == 0 && threadIdx.x == 0) {
        kf->proj_begin = proj_begin;
        kf->first = (unsigned) first;
        kf->n_proj = n_proj;
        kf->status = n_proj == 0 ? 1 : 0; // le_vins.cc:426-429 / 469
        kf->m = 0;
        kf->ins_t0 = t_first;
        kf->ins_t_last = t_last;
        kf->inv_mean_dt = inv_mean_dt;
    }

    // this slice's part of the per-interval table
    double R0[9], t0[3];
#pragma unroll
    for (int i = 0; i < 9; i++) R0[i] = sR0[i];
    t0[0] = st0[0];
    t0[1] = st0[1];
    t0[2] = st0[2];
    Interval* __restrict__ tab = a.table + kf->table_begin;
    const int per_slice = (nins + (int) n_slices - 1) / (int) n_slices;
    const int j_end = min(nins, (int) (slice + 1) * per_slice);
    for (int j = (int) slice * per_slice + (int) threadIdx.x; j < j_end; j += blockDim.x) {
        Interval e;
        double Rq[9], C[9], d[3];
        if (j == 0) { // out-of-window fallback: newest state, no interpolation (misc.cc:74-77)
            const InsSample s = w[nins - 1];
            quat_to_matrix(s.q, Rq);
            // (last, first): entry_brackets() is then true exactly outside the window (device_math.cuh)
            e.t_lo = nins >= 2 ? s.time : 0.0;
            e.t_hi = nins >= 2 ? w[0].time : 0.0;
            e.ts0 = 0;
            e.inv_dts = 0;
            e.k[0] = 1;
            e.k[1] = 0;
            e.k[2] = 0;
            e.theta = 0;
            d[0] = s.p[0] - t0[0];
            d[1] = s.p[1] - t0[1];
            d[2] = s.p[2] - t0[2];
            m3t_mul_v(R0, d, e.c0);
            e.c1[0] = e.c1[1] = e.c1[2] = 0;
        } else {
            const InsSample s0 = w[j - 1], s1 = w[j];
            double q1inv[4], dq[4];
            quat_inverse(s1.q, q1inv);
            quat_mul(q1inv, s0.q, dq); // misc.cc:85
            quat_to_angle_axis(dq, e.theta, e.k);
            double q0[4] = {s0.q[0], s0.q[1], s0.q[2], s0.q[3]};
            quat_normalize(q0); // the reference normalises the interpolated quaternion (misc.cc:94)
            quat_to_matrix(q0, Rq);
            e.t_lo = s0.time;
            e.t_hi = s1.time;
            e.ts0 = s0.state_time;
            e.inv_dts = 1.0 / (s1.state_time - s0.state_time);
            d[0] = s0.p[0] - t0[0];
            d[1] = s0.p[1] - t0[1];
            d[2] = s0.p[2] - t0[2];
            m3t_mul_v(R0, d, e.c0);
            d[0] = s1.p[0] - s0.p[0];
            d[1] = s1.p[1] - s0.p[1];
            d[2] = s1.p[2] - s0.p[2];
            m3t_mul_v(R0, d, e.c1);
        }
        m3t_mul_m(R0, Rq, C); // column-major
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int cI = 0; cI < 3; cI++) e.A[3 * r + cI] = C[3 * cI + r];
        // the one-compare bracket test needs positive, increasing times; the straight-line deskew a small angle
        const bool ordered = j == 0 ? (e.t_hi > 0 && e.t_lo > e.t_hi) : (e.t_lo > 0 && e.t_hi > e.t_lo);
        e.fast = (ordered && e.theta * e.theta < 2.44140625e-4) ? 1.0 : 0.0;
        fill_cull_record(e, a.ext, j == 0);
        tab[j] = e;
        if (j == 0) tab[nins] = e; // slot nins: the fallback again, as the successor of the last interval
    }

    // this slice's part of the chunk descriptors (k_deskew_bin's work list).  The interval hints come from the IMU times of
    // the window (input), not from the table (other slices are still writing it).
    {
        ChunkDesc* __restrict__ out = a.chunks + (size_t) blockIdx.x * a.chunks_per_kf;
        const unsigned long long table_begin = kf->table_begin;
        const unsigned cps = (a.chunks_per_kf + n_slices - 1) / n_slices;
        const unsigned c_begin = slice * cps, c_end = min(a.chunks_per_kf, c_begin + cps);
        // all first-point times of this thread's chunks first (independent loads, one memory round trip), then the probes;
        // chunks past the projected points only get an empty descriptor
        constexpr int ROUNDS = 8;
        for (unsigned base = c_begin; base < c_end; base += ROUNDS * 256u) {
            if ((unsigned long long) base * CHUNK_PTS >= n_proj) { // nothing live from here on
                ChunkDesc e;
                e.phys = 0;
                e.ent0 = 0;
                e.cnt = 0;
                for (unsigned c = base + threadIdx.x; c < c_end; c += 256u) out[c] = e;
                break;
            }
            double tp[ROUNDS];
            unsigned long long ph[ROUNDS];
#pragma unroll
            for (int i = 0; i < ROUNDS; i++) {
                const unsigned c = base + threadIdx.x + i * 256u;
                const unsigned long long off = (unsigned long long) c * CHUNK_PTS;
                const bool live = c < c_end && off < n_proj;
                ph[i] = live ? wrap(proj_begin + off, a.pts_cap) : 0ull;
                tp[i] = live ? a.pts[ph[i]].time : 0.0;
            }
#pragma unroll
            for (int i = 0; i < ROUNDS; i++) {
                const unsigned c = base + threadIdx.x + i * 256u;
                if (c < c_end)
                    out[c] = chunk_desc_from_time(c, n_proj, ph[i], tp[i], a.ext.td_b_l, w, table_begin, nins, t_first, t_last, inv_mean_dt);
            }
        }
    }

    // streaming: the pops are persistent (the deque is shared state).  Every block of the grid read the ring state at its
    // start, so the last one to get here may change it.
    if (a.ring) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned done = atomicAdd(&kf->ticket, 1u);
            if (done == gridDim.x * gridDim.y - 1) {
                a.ring->head = proj_begin;
                a.ring->count = count - first;
                kf->ticket = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_deskew_bin: persistent warps.  Every warp is its own pipeline: it walks the chunk list (128 consecutive points =
// 4 KB per chunk; warp g takes chunks g, g + W, g + 2W, ...) and lane 0 streams chunk it+2 HBM -> shared memory with the
// TMA unit (cp.async.bulk, 1-D, completion on the warp's own mbarrier) while the warp does the arithmetic of chunk it.
// With the points travel the two INS-interval entries the chunk can need (the one of its first point, found by
// k_chunk_desc, and its successor in the table), so that the per-point math never waits on global memory and the warps of a CTA
// never wait for each other.  Each thread handles 4 consecutive points; cell ids that the polynomial path cannot decide
// are queued per warp and settled 32 at a time by the exact path.
// ------------------------------------------------------------------------------------------------
#ifndef LV_GROUP
#define LV_GROUP 4
#endif
#ifndef LV_STAGES
#define LV_STAGES 3
#endif
#ifndef LV_AMBQ
#define LV_AMBQ 32
#endif
#ifndef LV_MINBLK
#define LV_MINBLK 1
#endif
#ifndef LV_CULL // 0: never try the certified frustum cull (A/B timing)
#define LV_CULL 1
#endif
#ifndef LV_KO // knock-out experiments (timing only, results are wrong): 1 no RED, 2 no deskew math, 4 no cell math, 8 loads only
#define LV_KO 0
#endif
constexpr int DESKEW_GROUP = LV_GROUP; // of which this many are in flight together (register pressure vs interleaving)
constexpr int DESKEW_WARPS = DESKEW_THREADS / 32;
constexpr int CHUNK_BYTES = CHUNK_PTS * 32;
constexpr int DESKEW_STAGES = LV_STAGES;
constexpr int STAGE_ENT_OFF = CHUNK_BYTES;                               // two interval entries (2 x 256 B)
constexpr int STAGE_BYTES = STAGE_ENT_OFF + 2 * (int) sizeof(Interval);
constexpr int AMBQ = LV_AMBQ; // undecided cell ids queued per warp
constexpr int WARP_BAR_OFF = DESKEW_STAGES * STAGE_BYTES; // one mbarrier per stage
constexpr int WARP_HDR_OFF = WARP_BAR_OFF + 32;           // points in the chunk of each stage
constexpr int WARP_Q_OFF = WARP_HDR_OFF + 16;
constexpr int WARP_QKF_OFF = WARP_Q_OFF + AMBQ * 16;
constexpr int WARP_SMEM = WARP_QKF_OFF + AMBQ * 4;
constexpr int DESKEW_SMEM = DESKEW_WARPS * WARP_SMEM; // 16 warps x 14 512 B = 232 192 B: one CTA per SM
constexpr int DESKEW_BLOCK = DESKEW_THREADS;
static_assert(DESKEW_STAGES <= 4, "barrier / header slots");
static_assert(STAGE_BYTES % 16 == 0 && WARP_SMEM % 16 == 0 && WARP_Q_OFF % 16 == 0, "TMA / vector alignment");
static_assert(LV_MINBLK * DESKEW_SMEM <= 227 * 1024, "CTAs per SM");

struct DeskewArgs {
    KfState* kfs; // first keyframe of this launch
    unsigned n_kf;          // keyframes in this launch
    unsigned chunks_per_kf; // ceil(max points per keyframe / CHUNK_PTS)
    const PointXYZIT* pts;
    unsigned long long pts_cap;
    const Interval* table;
    ChunkDesc* chunks;        // [n_kf][chunks_per_kf]
    unsigned long long* zbuf; // [n_kf][CELLS]
    PointXYZI* cloud_out;     // streaming only (n_kf == 1), nullable
    unsigned gen;             // z-buffer generation of this launch
    unsigned* cull_stat;      // nullable: {chunks walked, chunks on which the cull fired}, accumulated
    Extrinsics ext;
    Consts c;
};

// ---- mbarrier / TMA bulk-copy primitives (PTX) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completing `bytes` on the mbarrier (size and addresses multiples of 16 B)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same with an L2 eviction-priority hint: the point stream is read exactly once, so it should not push the
// z-buffers (re-used by every chunk of the keyframe and again by k_associate) out of L2
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar,
                                                  unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

__device__ __forceinline__ KfView kf_view(const KfState* __restrict__ kf, const Interval* table) {
    KfView kv;
    kv.tab = table + kf->table_begin;
    kv.n_ins = (int) kf->ins_count;
    kv.t0 = kf->ins_t0;
    kv.t_last = kf->ins_t_last;
    kv.inv_mean_dt = kf->inv_mean_dt;
    return kv;
}

template <int LIBM_MODE>
__device__ __noinline__ int cell_of_point_exact(float x, float y, float z, const Consts& c) {
    return cell_of_point<LIBM_MODE>(x, y, z, c);
}

// cell id: polynomial fast path, exact mirrored path next to rounding boundaries (bit-identical by construction)
template <int LIBM_MODE>
__device__ __forceinline__ int cell_id(float x, float y, float z, const Consts& c) {
    bool amb;
    int cell = cell_of_point_fast(x, y, z, c, amb);
    if (amb) cell = cell_of_point_exact<LIBM_MODE>(x, y, z, c);
    return cell;
}

// lane 0: start the copies of one chunk into a stage of the warp's ring
__device__ __forceinline__ void issue_chunk(const DeskewArgs& a, const ChunkDesc& d, unsigned char* stage, unsigned* hdr,
                                            unsigned long long* bar, unsigned long long policy) {
    const unsigned cnt = d.cnt;
    *hdr = cnt;
    if (cnt == 0) {
        mbar_arrive(bar); // release: the header is visible to whoever passes the barrier
        return;
    }
    constexpr unsigned ENT = 2u * (unsigned) sizeof(Interval); // entries ent0, ent0 + 1
    mbar_arrive_expect_tx(bar, cnt * 32u + ENT);
    tma_bulk_g2s(stage + STAGE_ENT_OFF, a.table + d.ent0, ENT, bar);
    const unsigned long long room = a.pts_cap - d.phys; // ring: the chunk may wrap around the end
    const unsigned first = room < cnt ? (unsigned) room : cnt;
    tma_bulk_g2s_hint(stage, a.pts + d.phys, first * 32u, bar, policy);
    if (first < cnt) tma_bulk_g2s_hint(stage + first * 32u, a.pts, (cnt - first) * 32u, bar, policy);
}

// settle the queued cell ids of a warp with the exact path (all lanes busy instead of one lane per rare case)
template <int LIBM_MODE>
__device__ __noinline__ void flush_undecided(const float4* q, const unsigned* qkf, unsigned n, unsigned long long* zbuf, unsigned gen,
                                             const Consts& c) {
    const unsigned lane = threadIdx.x & 31;
    __syncwarp();
    if (lane < n) {
        const float4 v = q[lane];
        const int cell = cell_of_point<LIBM_MODE>(v.x, v.y, v.z, c);
        const float range = sqrtf(v3f_sqnorm(v.x, v.y, v.z)); // tracking.cc:330
        if (cell >= 0 && (range < FLT_MAX)) atomicMin(zbuf + (size_t) qkf[lane] * CELLS + cell, key_make(gen, range, __float_as_uint(v.w)));
    }
    __syncwarp();
}

// Certified frustum cull (see Interval / fill_cull_record): true when the point is rejected by the FOV gate of
// tracking.cc:311-314 for certain, i.e. whatever its exact deskewed coordinates (X, Y, Z) are, given |(X,Y,Z) - p~| <= d:
//   z~ < -d                        =>  Z < 0
//   max(|x~|, |y~|) - |z~| > 2 d   =>  max(|X|, |Y|) > |Z|  =>  Z < 0 or fl32(|X| / |Z|) >= 1 > tan(45 deg) (0x1.fffffffffffffp-1)
// Any NaN makes both comparisons false (not culled).
struct CullRegs {
    float4 r0, r1, r2, r3; // M0..M3 | M4..M7 | M8 m0 m1 m2 | ca cb - -
};
__device__ __forceinline__ bool certainly_outside(const CullRegs& q, float x, float y, float z, float n2) {
    const float xt = fmaf(q.r0.z, z, fmaf(q.r0.y, y, fmaf(q.r0.x, x, q.r2.y)));
    const float yt = fmaf(q.r1.y, z, fmaf(q.r1.x, y, fmaf(q.r0.w, x, q.r2.z)));
    const float zt = fmaf(q.r2.x, z, fmaf(q.r1.w, y, fmaf(q.r1.z, x, q.r2.w)));
    float n;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(n) : "f"(n2)); // 2^-22 relative: covered by the 1 % the bound was inflated by
    const float d = fmaf(q.r3.x, n, q.r3.y);
    return (zt < -d) | (fmaxf(fabsf(xt), fabsf(yt)) - fabsf(zt) > 2.001f * d);
}

template <int LIBM_MODE, bool WRITE_CLOUD, bool CULL>
__global__ void __launch_bounds__(DESKEW_BLOCK, LV_MINBLK) k_deskew_bin(const __grid_constant__ DeskewArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + warp * WARP_SMEM;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(wbase + WARP_BAR_OFF);
    unsigned* hdr = reinterpret_cast<unsigned*>(wbase + WARP_HDR_OFF);
    float4* q_pt = reinterpret_cast<float4*>(wbase + WARP_Q_OFF);
    unsigned* q_kf = reinterpret_cast<unsigned*>(wbase + WARP_QKF_OFF);

    const unsigned W = gridDim.x * DESKEW_WARPS; // the warps of a CTA take neighbouring chunks of one keyframe
    const unsigned cpk = a.chunks_per_kf, total = a.n_kf * cpk;
    unsigned c = blockIdx.x * DESKEW_WARPS + warp;
    if (c >= total) return;
    unsigned kfi = c / cpk, rem = c - kfi * cpk; // keyframe and chunk-in-keyframe of chunk c, advanced incrementally
    const unsigned Wq = W / cpk, Wr = W - Wq * cpk;
    const unsigned long long policy = l2_policy_evict_first();

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < DESKEW_STAGES; s++) mbar_init(full + s, 1);
        mbar_fence_init();
#pragma unroll
        for (int i = 0; i < DESKEW_STAGES - 1; i++) {
            const unsigned cc = c + i * W;
            if (cc < total) {
                const uint4 r = __ldg(reinterpret_cast<const uint4*>(a.chunks + cc));
                ChunkDesc d;
                d.phys = (unsigned long long) r.x | ((unsigned long long) r.y << 32);
                d.ent0 = r.z;
                d.cnt = r.w;
                issue_chunk(a, d, wbase + i * STAGE_BYTES, hdr + i, full + i, policy);
            }
        }
    }
    // descriptor of the chunk that is issued next (one iteration ahead, so lane 0 never waits for it)
    uint4 dn = make_uint4(0, 0, 0, 0);
    if (c + (DESKEW_STAGES - 1) * W < total) dn = __ldg(reinterpret_cast<const uint4*>(a.chunks + c + (DESKEW_STAGES - 1) * W));

    unsigned s = 0, ph = 0, s_fill = DESKEW_STAGES - 1, qn = 0;
    // The cull only pays where whole warps fall outside the image cone (a 360-degree scanner: three quarters of the
    // chunks); a sensor that looks where the camera looks never culls anything.  After 8 chunks without a single culled
    // point the warp tries only every 16th chunk.
    unsigned cull_miss = 0, iter = 0, stat_hit = 0;
    for (; c < total; c += W, iter++) {
        // stage s_fill was read during the previous iteration: every lane is past those reads here
        __syncwarp();
        if (lane == 0 && c + (DESKEW_STAGES - 1) * W < total) {
            ChunkDesc d;
            d.phys = (unsigned long long) dn.x | ((unsigned long long) dn.y << 32);
            d.ent0 = dn.z;
            d.cnt = dn.w;
            issue_chunk(a, d, wbase + s_fill * STAGE_BYTES, hdr + s_fill, full + s_fill, policy);
        }
        if (c + DESKEW_STAGES * W < total) dn = __ldg(reinterpret_cast<const uint4*>(a.chunks + c + DESKEW_STAGES * W));
        const unsigned char* st = wbase + s * STAGE_BYTES;
        mbar_wait(full + s, ph);
        const unsigned cnt = hdr[s];

        if (cnt) {
            const float4* __restrict__ sp = reinterpret_cast<const float4*>(st);
            const unsigned off = rem * CHUNK_PTS;
            const bool cull_try = LV_CULL && CULL && !WRITE_CLOUD && (cull_miss < 8 || (iter & 15) == 0);
            bool culled_any = false;
            // 4 consecutive points per thread (they almost always share one INS interval).  Lane l visits them in the order
            // (j + l) & 3: with the plain order all lanes of a quarter-warp would hit the same 4 banks (128-byte lane stride,
            // 8-way conflict on every LDS.128); rotated, the conflict is 2-way.
            const unsigned i0 = lane * DESKEW_PPT;
#pragma unroll 1
            for (int g0 = 0; g0 < DESKEW_PPT; g0 += DESKEW_GROUP) {
            unsigned pi[DESKEW_GROUP];
#pragma unroll
            for (int j = 0; j < DESKEW_GROUP; j++) pi[j] = i0 + ((g0 + j + lane) & (DESKEW_PPT - 1));
            float px[DESKEW_GROUP], py[DESKEW_GROUP], pz[DESKEW_GROUP], pn2[DESKEW_GROUP];
            double tt[DESKEW_GROUP];
            bool kept[DESKEW_GROUP];
            bool any_kept = false;
#pragma unroll
            for (int j = 0; j < DESKEW_GROUP; j++) {
                const float4 xyz = sp[2 * pi[j]];
                const double tm = reinterpret_cast<const double2*>(sp)[2 * pi[j] + 1].y;
                px[j] = xyz.x;
                py[j] = xyz.y;
                pz[j] = xyz.z;
                tt[j] = tm + a.ext.td_b_l; // le_vins.cc:455
                pn2[j] = v3f_sqnorm(xyz.x, xyz.y, xyz.z);
                kept[j] = (pi[j] < cnt) && !((pn2[j] > a.c.n2_far) || (pn2[j] < a.c.n2_near)); // le_vins.cc:447-451
                any_kept |= kept[j];
            }
            float ox[DESKEW_GROUP], oy[DESKEW_GROUP], oz[DESKEW_GROUP];
#if LV_KO & 8
            {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < DESKEW_GROUP; j++) acc += px[j] + py[j] + pz[j] + (float) tt[j];
                if (acc == 12345.678f) atomicMin(a.zbuf, (unsigned long long) __float_as_uint(acc));
                continue;
            }
#endif
#if LV_KO & 2
#pragma unroll
            for (int j = 0; j < DESKEW_GROUP; j++) {
                ox[j] = px[j] + (float) tt[j];
                oy[j] = py[j];
                oz[j] = pz[j];
            }
            if (false) {
#else
            if (any_kept) {
#endif
                // which of the two staged entries brackets every kept point of this thread (checked, not assumed)
                const unsigned char* E0 = st + STAGE_ENT_OFF;
                const unsigned char* E1 = E0 + sizeof(Interval);
                auto usable = [](const unsigned char* e) { return reinterpret_cast<const Interval*>(e)->fast != 0.0; };
                const double2 T0 = *reinterpret_cast<const double2*>(E0);
                const unsigned long long lo0 = (unsigned long long) __double_as_longlong(T0.x);
                const unsigned long long w0 = usable(E0) ? (unsigned long long) __double_as_longlong(T0.y) - lo0 : 0ull;
                bool in0 = true;
#pragma unroll
                for (int j = 0; j < DESKEW_GROUP; j++) in0 &= !kept[j] || key_brackets(lo0, w0, tt[j]);
                const unsigned char* ep = E0;
                bool fast = in0;
                unsigned long long lo1 = 0, w1 = 0;
                if (!in0) {
                    const double2 T1 = *reinterpret_cast<const double2*>(E1);
                    lo1 = (unsigned long long) __double_as_longlong(T1.x);
                    w1 = usable(E1) ? (unsigned long long) __double_as_longlong(T1.y) - lo1 : 0ull;
                    bool in1 = true;
#pragma unroll
                    for (int j = 0; j < DESKEW_GROUP; j++) in1 &= !kept[j] || key_brackets(lo1, w1, tt[j]);
                    ep = E1;
                    fast = in1;
                }
                if (fast) {
                    if (cull_try) {
                        // points that cannot land inside the image cone need no coordinates (and leave no trace)
                        const float4* cr = reinterpret_cast<const float4*>(ep + offsetof(Interval, M));
                        CullRegs q;
                        q.r3 = cr[3];
                        if (q.r3.x >= 0.f) {
                            q.r0 = cr[0];
                            q.r1 = cr[1];
                            q.r2 = cr[2];
                            any_kept = false;
#pragma unroll
                            for (int j = 0; j < DESKEW_GROUP; j++) {
                                const bool out = kept[j] && certainly_outside(q, px[j], py[j], pz[j], pn2[j]);
                                culled_any |= out;
                                kept[j] = kept[j] && !out;
                                any_kept |= kept[j];
                            }
                        }
                    }
                    if (any_kept) {
                        // common case: one entry for the group, straight-line code, the points interleave
                        EntryRegs e;
                        load_entry_shared(reinterpret_cast<const Interval*>(ep), e);
#pragma unroll
                        for (int j = 0; j < DESKEW_GROUP; j++) deskew_core<true>(px[j], py[j], pz[j], tt[j], e, a.ext, ox[j], oy[j], oz[j]);
                    }
                } else {
                    // the thread's points straddle intervals, or need an entry that was not staged (window edges, large
                    // rotation, unordered times): one entry per point
#pragma unroll
                    for (int j = 0; j < DESKEW_GROUP; j++)
                        if (kept[j]) {
                            const unsigned char* es = key_brackets(lo0, w0, tt[j]) ? E0 : key_brackets(lo1, w1, tt[j]) ? E1 : nullptr;
                            const Interval* epj = es ? reinterpret_cast<const Interval*>(es) : find_entry_global(a.kfs + kfi, a.table, tt[j]);
                            EntryRegs e;
                            load_entry_shared(epj, e); // generic pointer: shared or global
                            float x, y, z;
                            deskew_core<false>(px[j], py[j], pz[j], tt[j], e, a.ext, x, y, z);
                            ox[j] = x;
                            oy[j] = y;
                            oz[j] = z;
                        }
                }
            }
            if (WRITE_CLOUD) {
#pragma unroll
                for (int j = 0; j < DESKEW_GROUP; j++)
                    if (pi[j] < cnt) {
                        float4* out = reinterpret_cast<float4*>(a.cloud_out + off + pi[j]);
                        out[0] = kept[j] ? make_float4(ox[j], oy[j], oz[j], 1.0f) : make_float4(0.f, 0.f, 0.f, 1.0f); // resize() default (le_vins.cc:438)
                        out[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
            }
            // nothing left in this warp (every point gated or culled): no cell ids, no keys
            if (!__any_sync(0xffffffffu, any_kept)) continue;
            // stage 2: cell ids (a default point lands at col = -150 and is dropped at tracking.cc:326-328)
            unsigned long long key[DESKEW_GROUP];
            int cell[DESKEW_GROUP];
            bool amb_any = false;
            bool amb[DESKEW_GROUP];
#pragma unroll
            for (int j = 0; j < DESKEW_GROUP; j++) {
#if LV_KO & 4
                cell[j] = (int) (__float_as_uint(ox[j] + oy[j] * oz[j]) % 90000u);
                amb[j] = false;
#else
                cell[j] = cell_of_point_fast(ox[j], oy[j], oz[j], a.c, amb[j]);
#endif
                amb[j] = amb[j] && kept[j];
                amb_any |= amb[j];
            }
#pragma unroll
            for (int j = 0; j < DESKEW_GROUP; j++) {
                const float range = sqrtf(v3f_sqnorm(ox[j], oy[j], oz[j])); // tracking.cc:330
                if (!kept[j] || !(range < FLT_MAX)) cell[j] = -1;          // tracking.cc:332 against the FLT_MAX initial value
                key[j] = key_make(a.gen, range, off + pi[j]);
            }
            // z-buffer: fire-and-forget RED.MIN.64, one per in-grid point (no probe: a probe costs the L1/L2 path as much
            // as the reduction itself and would put an L2 round trip on the critical path)
            unsigned long long* __restrict__ zb = a.zbuf + (size_t) kfi * CELLS;
#pragma unroll
            for (int j = 0; j < DESKEW_GROUP; j++)
#if LV_KO & 1
                if (cell[j] == 123456789 && key[j] == 42) atomicMin(zb, key[j]);
#else
                if (cell[j] >= 0) atomicMin(zb + cell[j], key[j]);
#endif
            // cell ids next to a rounding boundary: queue them, the exact path settles 32 at a time
            if (__any_sync(0xffffffffu, amb_any)) {
#pragma unroll
                for (int j = 0; j < DESKEW_GROUP; j++) {
                    const unsigned m = __ballot_sync(0xffffffffu, amb[j]);
                    if (m) {
                        const unsigned n = __popc(m);
                        if (qn + n > AMBQ) {
                            flush_undecided<LIBM_MODE>(q_pt, q_kf, qn, a.zbuf, a.gen, a.c);
                            qn = 0;
                        }
                        if (amb[j]) {
                            const unsigned pos = qn + __popc(m & ((1u << lane) - 1u));
                            q_pt[pos] = make_float4(ox[j], oy[j], oz[j], __uint_as_float(off + pi[j]));
                            q_kf[pos] = kfi;
                        }
                        qn += n;
                    }
                }
            }
            } // group
            if (cull_try) {
                const bool hit = __any_sync(0xffffffffu, culled_any);
                cull_miss = hit ? 0u : cull_miss + 1u;
                stat_hit += hit ? 1u : 0u;
            }
        }
        s_fill = s;
        if (++s == DESKEW_STAGES) {
            s = 0;
            ph ^= 1;
        }
        rem += Wr;
        kfi += Wq;
        if (rem >= cpk) {
            rem -= cpk;
            kfi++;
        }
    }
    if (qn) flush_undecided<LIBM_MODE>(q_pt, q_kf, qn, a.zbuf, a.gen, a.c);
    // how often the cull fired (chunks): the host picks the instantiation without the cull for sensors that never leave the
    // image cone, and probes again now and then
    if (CULL && a.cull_stat && lane == 0 && iter) {
        atomicAdd(a.cull_stat, iter); // chunks this warp walked (the cull was tried on all of them unless it had backed off)
        atomicAdd(a.cull_stat + 1, stat_hit);
    }
}


// explicit-cloud binning for levins_build_range_image(cloud != NULL)
template <int LIBM_MODE>
__global__ void __launch_bounds__(256) k_bin_cloud(const PointXYZI* __restrict__ cloud, unsigned n, unsigned long long* zbuf,
                                                   unsigned gen, const __grid_constant__ Consts c) {
    const unsigned k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const float4 p = __ldg(reinterpret_cast<const float4*>(cloud + k));
    const int cell = cell_id<LIBM_MODE>(p.x, p.y, p.z, c);
    if (cell < 0) return;
    const float range = sqrtf(v3f_sqnorm(p.x, p.y, p.z));
    if (!(range < FLT_MAX)) return;
    const unsigned long long key = key_make(gen, range, k);
    if (key < __ldcg(zbuf + cell)) atomicMin(zbuf + cell, key);
}

// unit vector + range of a camera-frame point (tracking.cc:346-351)
__device__ __forceinline__ float4 unit_and_range(float x, float y, float z) {
    const float n2 = v3f_sqnorm(x, y, z);
    const float r = sqrtf(n2);
    if (n2 > 0.f) {
        const float nn = sqrtf(n2);
        return make_float4(x / nn, y / nn, z / nn, r);
    }
    return make_float4(x, y, z, r);
}

// ------------------------------------------------------------------------------------------------
// k_finalize (stage-wise entry points only): one thread per cell.
// cells[] = (unit x, unit y, unit z, range), range = -1 marks an empty cell.
// ------------------------------------------------------------------------------------------------
struct FinalizeArgs {
    KfState* kfs;
    const PointXYZIT* pts;
    unsigned long long pts_cap;
    const Interval* table;
    const PointXYZI* cloud; // FROM_CLOUD source
    const unsigned long long* zbuf;
    float4* cells;   // [gridDim.y][CELLS]
    int32_t* winner; // nullable, [CELLS] (single keyframe)
    unsigned gen;
    Extrinsics ext;
    Consts c;
};

template <bool FROM_CLOUD>
__global__ void __launch_bounds__(FINALIZE_THREADS) k_finalize(const __grid_constant__ FinalizeArgs a) {
    KfState* __restrict__ kf = a.kfs + blockIdx.y;
    const int cell = blockIdx.x * FINALIZE_THREADS + threadIdx.x;
    bool filled = false;
    if (cell < CELLS) {
        const unsigned long long key = a.zbuf[(size_t) blockIdx.y * CELLS + cell];
        float4 out = make_float4(0.f, 0.f, 0.f, -1.0f);
        int win = -1;
        if (key_live(key, a.gen)) {
            const unsigned k = key_index(key);
            float x, y, z;
            if (FROM_CLOUD) {
                const float4 p = __ldg(reinterpret_cast<const float4*>(a.cloud + k));
                x = p.x;
                y = p.y;
                z = p.z;
            } else { // recompute the winner's deskew: same function, same bits as in k_deskew_bin
                const KfView kv = kf_view(kf, a.table);
                const float4* __restrict__ p4 = reinterpret_cast<const float4*>(a.pts + wrap(kf->proj_begin + k, a.pts_cap));
                const float4 xyz = __ldg(p4);
                const double tm = __ldg(reinterpret_cast<const double2*>(p4 + 1)).y;
                EntryRegs e;
                int cur = -1;
                deskew_point(xyz.x, xyz.y, xyz.z, tm, kv, a.ext, a.c, e, cur, x, y, z);
            }
            out = unit_and_range(x, y, z);
            win = (int) k;
            filled = true;
        }
        a.cells[(size_t) blockIdx.y * CELLS + cell] = out;
        if (a.winner) a.winner[cell] = win;
    }
    const unsigned cnt = __syncthreads_count(filled); // M = number of non-empty cells (tracking.cc:356-358)
    if (threadIdx.x == 0 && cnt) atomicAdd(&kf->m, cnt);
}

// M for the fused paths: counts the live cells of [blockIdx.y]'s z-buffer slice owned by this block
__device__ __forceinline__ void count_live_cells(const unsigned long long* __restrict__ zb, unsigned gen, KfState* kf) {
    const int chunk = (CELLS + gridDim.x - 1) / gridDim.x;
    const int begin = blockIdx.x * chunk, end = min(begin + chunk, CELLS);
    unsigned c = 0;
    const int stride = blockDim.x;
    int i = begin + threadIdx.x;
    for (; i + 3 * stride < end; i += 4 * stride) { // 4 independent loads in flight
        const unsigned long long k0 = __ldg(zb + i), k1 = __ldg(zb + i + stride), k2 = __ldg(zb + i + 2 * stride),
                                 k3 = __ldg(zb + i + 3 * stride);
        c += key_live(k0, gen) + key_live(k1, gen) + key_live(k2, gen) + key_live(k3, gen);
    }
    for (; i < end; i += stride) c += key_live(__ldg(zb + i), gen);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&kf->m, c);
}

__global__ void __launch_bounds__(256) k_count_cells(KfState* kfs, const unsigned long long* __restrict__ zbuf, unsigned gen) {
    count_live_cells(zbuf + (size_t) blockIdx.y * CELLS, gen, kfs + blockIdx.y);
}

// ------------------------------------------------------------------------------------------------
// k_compact: row-major compaction (tracking.cc:341-351) -> cell-to-compact-index map + compacted cloud.
// One block of 1024 threads per keyframe; only the stage-wise entry points need it (indices are an
// internal quantity of the reference; the fused path orders neighbours by cell id, which is the same order).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_compact(const float4* __restrict__ cells, int32_t* __restrict__ c2c,
                                                  PointXYZI* __restrict__ unit_sphere) {
    constexpr int PER = (CELLS + 1023) / 1024; // 88
    __shared__ unsigned warp_sums[32];
    const int t = threadIdx.x;
    const int begin = t * PER;
    unsigned cnt = 0;
    for (int i = 0; i < PER; i++) {
        const int cidx = begin + i;
        if (cidx < CELLS && cells[cidx].w >= 0.f) cnt++;
    }
    // block exclusive scan of cnt
    unsigned incl = cnt;
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned v = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        warp_sums[lane] = v;
    }
    __syncthreads();
    unsigned pos = incl - cnt + (warp > 0 ? warp_sums[warp - 1] : 0);
    for (int i = 0; i < PER; i++) {
        const int cidx = begin + i;
        if (cidx >= CELLS) break;
        const float4 v = cells[cidx];
        if (v.w >= 0.f) {
            c2c[cidx] = (int32_t) pos;
            if (unit_sphere) {
                float4* o = reinterpret_cast<float4*>(unit_sphere + pos);
                o[0] = make_float4(v.x, v.y, v.z, 1.0f);
                o[1] = make_float4(v.w, 0.f, 0.f, 0.f);
            }
            pos++;
        } else {
            c2c[cidx] = -1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_associate: one warp per feature.
//   FUSED = false: candidates come from the finalised cell array (stage-wise API).
//   FUSED = true : candidates come straight from the z-buffer: the winner of each window cell is deskewed on the fly
//                  (same function, same bits as in k_deskew_bin), so the batch / keyframe path needs no k_finalize,
//                  no 1.44 MB cell array per keyframe and only deskews the cells some feature actually looks at.
// ------------------------------------------------------------------------------------------------
struct AssocArgs {
    KfState* kfs;
    const float* queries; // pool, 3 floats per query
    const float4* cells;  // !FUSED: [gridDim.y][CELLS]
    const int32_t* c2c;   // nullable (single keyframe): cell -> compact index
    // FUSED sources
    const unsigned long long* zbuf; // [gridDim.y][CELLS]
    const PointXYZIT* pts;
    unsigned long long pts_cap;
    const Interval* table;
    unsigned gen;
    int count_cells; // FUSED: also count the live cells of the z-buffer (M) in this launch
    Extrinsics ext;
    // outputs
    // feature side on the device (levins_associate_pixels): queries are made here from tracked pixels
    const float* pixels; // nullable; 2 floats per query, same indexing as `queries`
    float* queries_out;  // the f32 rays, stored for k_plane (the query pool itself)
    float* undist_out;   // nullable: the undistorted pixels
    int undistort;       // run Camera::undistortPoints first (camera.cc:72-74)
    CameraModel cam;
    float4* nn;       // 5 per query: neighbours (unit vector + range) for k_plane
    uint2* nn_meta;   // (found, bits of the 5th squared distance)
    int32_t* knn_idx; // nullable, 5 per query
    float* knn_d2;    // nullable
    Consts c;
};

template <bool FUSED>
__global__ void __launch_bounds__(ASSOC_WARPS * 32, LV_ASSOC_MINBLK) k_associate(const __grid_constant__ AssocArgs a) {
    KfState* __restrict__ kf = a.kfs + blockIdx.y;
    const int lane = threadIdx.x & 31;
    const unsigned f = blockIdx.x * ASSOC_WARPS + (threadIdx.x >> 5);
    if (f >= kf->q_count) {
        if (FUSED && a.count_cells) count_live_cells(a.zbuf + (size_t) blockIdx.y * CELLS, a.gen, kf);
        return;
    }
    const size_t qi = kf->q_begin + f;
    float qx, qy, qz;
    if (a.pixels) { // tracking.cc:381-395: undistort, pixel2unitcam, cast to float (every lane computes the same ray)
        float pu = __ldg(a.pixels + 2 * qi), pv = __ldg(a.pixels + 2 * qi + 1);
        if (a.undistort) undistort_pixel(a.cam, pu, pv);
        pixel_to_unit_ray(a.cam, pu, pv, qx, qy, qz);
        if (lane == 0) {
            a.queries_out[3 * qi] = qx;
            a.queries_out[3 * qi + 1] = qy;
            a.queries_out[3 * qi + 2] = qz;
            if (a.undist_out) {
                a.undist_out[2 * qi] = pu;
                a.undist_out[2 * qi + 1] = pv;
            }
        }
    } else {
        qx = __ldg(a.queries + 3 * qi);
        qy = __ldg(a.queries + 3 * qi + 1);
        qz = __ldg(a.queries + 3 * qi + 2);
    }

    // The query's position in cell units.  Precision is irrelevant: the window and the mask below carry half a cell
    // of slack, and every candidate they keep is measured with the reference's exact f32 arithmetic afterwards.
    float rq, cq, ael;
    const float inv_ang = a.c.inv_angle_f, off = (float) a.c.half_offset;
    if (qz > 0.f && fabsf(qx) <= qz && fabsf(qy) <= qz) { // inside the 90-degree cone: both arguments lie in [-1, 1]
        const float el = atan_poly(qy * rsqrtf(fmaf(qx, qx, qz * qz)));
        rq = fmaf(el, inv_ang, 150.0f);
        cq = fmaf(-atan_poly(__fdividef(qx, qz)), inv_ang, 150.0f);
        ael = fabsf(el);
    } else {
        const float el = atan2f(qy, sqrtf(qx * qx + qz * qz));
        const float az = atan2f(qz, qx);
        rq = (el + off) * inv_ang;
        cq = (az - off) * inv_ang;
        ael = fabsf(el);
    }
    const int rowq = (int) rintf(rq), colq = (int) rintf(cq);
    // Only points within the acceptance radius T = 3 sin(0.3 deg) (chord; 3.00002 cells of arc) can influence a result
    // (tracking.cc:573).  A point of cell (r, c) lies within +-0.5 cell of (r, c) in (row, col) angle, and by the
    // haversine formula its arc distance g to the query obeys  g^2 >= (1 - 3e-4) (dEl^2 + cos(el_q) cos(el_p) dAz^2),
    // |el_p| <= |el_q| + 0.91 deg.  Cells that cannot contain such a point are skipped (about 60 % of the window).
    const float cmin = fmaxf(0.f, __cosf(ael) * __cosf(fminf(ael + 0.016f, 1.5707963f))) * 0.999f;

    // ---- candidates: scan the 9 x 13 window, keep live cells inside the mask, compact them across the warp ----
    __shared__ unsigned long long s_key[ASSOC_WARPS][128];
    __shared__ int s_cid[ASSOC_WARPS][128];
    const int wib = threadIdx.x >> 5;
    const unsigned long long* __restrict__ zb = FUSED ? a.zbuf + (size_t) blockIdx.y * CELLS : nullptr;
    const float4* __restrict__ cells = FUSED ? nullptr : a.cells + (size_t) blockIdx.y * CELLS;
    int count = 0;
    bool inmask[4];
    unsigned long long wkey[4];
    int wcell[4];
#pragma unroll
    for (int itc = 0; itc < 4; itc++) { // all loads of the window first: one L2 round trip, not four
        const int j = lane + 32 * itc;
        const int r = rowq - WIN_ROWS + j / WIN_W;
        const int cc = colq - WIN_COLS + j % WIN_W;
        inmask[itc] = false;
        wkey[itc] = 0;
        wcell[itc] = r * IMG + cc;
        if (j < WIN_CELLS && r >= 0 && r < IMG && cc >= 0 && cc < IMG) {
            const float dr = fmaxf(0.f, fabsf((float) r - rq) - 0.51f);
            const float dc = fmaxf(0.f, fabsf((float) cc - cq) - 0.51f);
            inmask[itc] = dr * dr + cmin * dc * dc <= 9.0601f; // 3.01^2
            if (inmask[itc]) {
                if (FUSED)
                    wkey[itc] = __ldg(zb + wcell[itc]);
                else
                    wkey[itc] = __float_as_uint(__ldg(&cells[wcell[itc]].w)); // range, -1 (sign bit set) for empty cells
            }
        }
    }
#pragma unroll
    for (int itc = 0; itc < 4; itc++) {
        const bool live = inmask[itc] && (FUSED ? key_live(wkey[itc], a.gen) : !((unsigned) wkey[itc] >> 31));
        const unsigned bal = __ballot_sync(0xffffffffu, live);
        if (live) {
            const int pos = count + __popc(bal & ((1u << lane) - 1));
            s_key[wib][pos] = wkey[itc];
            s_cid[wib][pos] = wcell[itc];
        }
        count += __popc(bal);
    }
    __syncwarp();

    float4 cand[4];
    int cid[4];
    unsigned dbits[4];
#pragma unroll
    for (int itc = 0; itc < 4; itc++) {
        cand[itc] = make_float4(0.f, 0.f, 0.f, -1.0f);
        cid[itc] = 0x7fffffff;
        dbits[itc] = 0x7f800000u; // +inf
    }
    {
        KfView kv{};
        unsigned long long pbegin = 0;
        if (FUSED) {
            kv = kf_view(kf, a.table);
            pbegin = kf->proj_begin;
        }
        EntryRegs e;
        int cur = -1;
#pragma unroll
        for (int itc = 0; itc < 4; itc++) {
            if (32 * itc < count) { // warp-uniform: rounds beyond the candidate count cost nothing
                const int ci = lane + 32 * itc;
                if (ci < count) {
                    cid[itc] = s_cid[wib][ci];
                    if (FUSED) {
                        // the winner of this cell, deskewed on the fly: same function, same bits as in k_deskew_bin
                        const float4* __restrict__ p4 =
                            reinterpret_cast<const float4*>(a.pts + wrap(pbegin + key_index(s_key[wib][ci]), a.pts_cap));
                        const float4 xyz = __ldg(p4);
                        const double tm = __ldg(reinterpret_cast<const double2*>(p4 + 1)).y;
                        float x, y, z;
                        deskew_point(xyz.x, xyz.y, xyz.z, tm, kv, a.ext, a.c, e, cur, x, y, z);
                        cand[itc] = unit_and_range(x, y, z); // tracking.cc:346-351
                    } else {
                        cand[itc] = __ldg(cells + cid[itc]);
                    }
                    // FLANN L2_Simple<float>: result = 0; result += diff*diff, x then y then z
                    const float dx = qx - cand[itc].x, dy = qy - cand[itc].y, dz = qz - cand[itc].z;
                    float d = dx * dx;
                    d = d + dy * dy;
                    d = d + dz * dz;
                    dbits[itc] = __float_as_uint(d);
                }
            }
        }
    }

    // ---- top-5 by (d2, cell id): cell ids are row-major, i.e. the order of the reference's compacted indices ----
    unsigned nn_d[5];
    int nn_c[5];
    float4 nn_p[5];
    int found = 0;
    auto select5 = [&](auto nslots_tag) {
        constexpr int NS = decltype(nslots_tag)::value;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            unsigned bd = dbits[0];
            int bc = cid[0];
            float4 bp = cand[0];
#pragma unroll
            for (int itc = 1; itc < NS; itc++)
                if (dbits[itc] < bd || (dbits[itc] == bd && cid[itc] < bc)) {
                    bd = dbits[itc];
                    bc = cid[itc];
                    bp = cand[itc];
                }
            const unsigned m1 = __reduce_min_sync(0xffffffffu, bd);
            const unsigned m2 = __reduce_min_sync(0xffffffffu, bd == m1 ? (unsigned) bc : 0xffffffffu);
            const bool mine = (bd == m1) && ((unsigned) bc == m2);
            const int owner = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
            nn_d[r] = m1;
            nn_c[r] = (int) m2;
            nn_p[r].x = __shfl_sync(0xffffffffu, bp.x, owner);
            nn_p[r].y = __shfl_sync(0xffffffffu, bp.y, owner);
            nn_p[r].z = __shfl_sync(0xffffffffu, bp.z, owner);
            nn_p[r].w = __shfl_sync(0xffffffffu, bp.w, owner);
            if (m1 != 0x7f800000u) found++;
            if (mine) { // this lane owns the winner: retire it
#pragma unroll
                for (int itc = 0; itc < NS; itc++)
                    if (dbits[itc] == m1 && cid[itc] == (int) m2) {
                        dbits[itc] = 0x7f800000u;
                        cid[itc] = 0x7fffffff;
                    }
            }
        }
    };
    if (count <= 32) // the usual case: at most one candidate per lane (warp-uniform branch)
        select5(std::integral_constant<int, 1>());
    else
        select5(std::integral_constant<int, 4>());

    // hand the 5 neighbours to k_plane (one THREAD per feature there: the 5x3 QR runs at full lane efficiency
    // instead of 32 lanes repeating one feature's solve)
    if (lane < 5) {
        float4 pp = nn_p[0];
        unsigned d = nn_d[0];
        int cI = nn_c[0];
#pragma unroll
        for (int k = 1; k < 5; k++)
            if (lane == k) {
                pp = nn_p[k];
                d = nn_d[k];
                cI = nn_c[k];
            }
        a.nn[5 * qi + lane] = pp;
        if (lane == 4) a.nn_meta[qi] = make_uint2((unsigned) found, d);
        if (a.knn_idx) {
            const bool ok = d != 0x7f800000u;
            a.knn_idx[5 * qi + lane] = ok ? (a.c2c ? a.c2c[cI] : cI) : -1;
            if (a.knn_d2) a.knn_d2[5 * qi + lane] = ok ? __uint_as_float(d) : FLT_MAX;
        }
    }
    // M (number of non-empty cells), piggy-backed on this launch: every block counts its slice of the z-buffer
    if (FUSED && a.count_cells) count_live_cells(a.zbuf + (size_t) blockIdx.y * CELLS, a.gen, kf);
}

// ------------------------------------------------------------------------------------------------
// k_plane: one thread per feature.  Threshold on the 5th distance, least-squares plane through the 5 metric points
// (Eigen ColPivHouseholderQR restated), the three validity checks and the ray-plane depth (tracking.cc:573-631),
// in the reference's operation order (no contraction in this translation unit).
// ------------------------------------------------------------------------------------------------
struct PlaneArgs {
    const float* queries;
    const float4* nn;     // 5 per query: unit vector + range of the neighbours, ascending (d2, index)
    const uint2* nn_meta; // (number found, bits of the 5th squared distance)
    unsigned long long q_begin, q_count;
    int8_t* type;
    float4* xyz_depth;
    int32_t* flags;
    Consts c;
    PeerOut peer; // peer mode: the results also go straight into every rank's gathered block (exchange.cuh)
    // feature side on the device: the drivers' accept gate and cam2world (tracking.cc:431-439, 539-545)
    // streaming keyframe: the final KfState and RingState are copied next to the results, so that one device->host copy
    // brings everything back (k_associate, which counts M, has completed when this kernel runs)
    const KfState* kf_src;
    KfState* kf_out;
    const RingState* ring_src;
    RingState* ring_out;
    unsigned char* accepted; // nullable: type != NONE && isGoodDepth(depth, 1)
    double* pw;              // nullable: 3 per query, world point of accepted features (0 otherwise)
    double poseR[9], poset[3]; // frame_ref_->pose(), R column-major
};

__global__ void __launch_bounds__(128) k_plane(const __grid_constant__ PlaneArgs a) {
    const unsigned long long i = (unsigned long long) blockIdx.x * 128 + threadIdx.x;
    if (i == 0 && a.kf_out) {
        *a.kf_out = *a.kf_src;
        if (a.ring_out) *a.ring_out = *a.ring_src;
    }
    if (i >= a.q_count) return;
    const size_t qi = a.q_begin + i;
    const float qx = __ldg(a.queries + 3 * qi), qy = __ldg(a.queries + 3 * qi + 1), qz = __ldg(a.queries + 3 * qi + 2);
    const uint2 meta = __ldg(a.nn_meta + qi);
    int type = -1, flags = 0;
    float ox = qx, oy = qy, oz = qz, od = 0.f;
    if (meta.x < 5) {
        flags = 1;
    } else if ((double) __uint_as_float(meta.y) > a.c.thr2) { // tracking.cc:573-575
        flags = 2;
    } else {
        double P[5][3];
        float rng0 = 0.f, rng4 = 0.f;
#pragma unroll
        for (int k = 0; k < 5; k++) { // tracking.cc:577-581
            const float4 v = __ldg(a.nn + 5 * qi + k);
            const double rr = (double) v.w;
            P[k][0] = (double) v.x * rr;
            P[k][1] = (double) v.y * rr;
            P[k][2] = (double) v.z * rr;
            if (k == 0) rng0 = v.w;
            if (k == 4) rng4 = v.w;
        }
        double mx[3];
        colpiv_qr_solve_5x3(P, mx);                            // tracking.cc:584-589
        const double norm_inverse = 1.0 / sqrt(v3_dot(mx, mx)); // :592
        {
            const double z = v3_dot(mx, mx); // :593 normalize()
            if (z > 0) {
                const double s = sqrt(z);
                mx[0] /= s;
                mx[1] /= s;
                mx[2] /= s;
            }
        }
        bool isvalid = true;
#pragma unroll
        for (int k = 0; k < 5; k++) { // :596-605 (the early break only skips tests whose outcome cannot matter)
            const double sm = fabs(v3_dot(mx, P[k]) + norm_inverse);
            if (isvalid && sm > a.c.plane_residual) {
                isvalid = false;
                flags |= 4;
            }
        }
        const float depth_diff = fabsf(rng0 - rng4); // :607-608
        if ((double) depth_diff > a.c.range_spread) {
            isvalid = false;
            flags |= 8;
        }
        if (isvalid) {
            const float vr0 = (float) P[4][0], vr1 = (float) P[4][1], vr2 = (float) P[4][2];
            const float vn0 = (float) mx[0], vn1 = (float) mx[1], vn2 = (float) mx[2];
            const float num = vn0 * vr0 + (vn1 * vr1 + vn2 * vr2); // Vector3f dot: e0 + (e1 + e2)
            const float den = vn0 * qx + (vn1 * qy + vn2 * qz);
            const float t = num / den; // :618
            if (isnan(t) || ((double) t < a.c.nearest) || ((double) t > a.c.farthest)) {
                flags |= 16; // :621-623
            } else {
                ox = qx * t; // :626
                oy = qy * t;
                oz = qz * t;
                od = oz; // :629
                type = 1;
            }
        }
    }
    a.type[qi] = (int8_t) type;
    a.xyz_depth[qi] = make_float4(ox, oy, oz, od);
    a.flags[qi] = flags;
    if (a.accepted) {
        const double depth = (double) od; // tracking.cc:430: double depth = point.intensity
        const bool ok = (type != -1) && (depth > a.c.nearest) && (depth < a.c.farthest * 1.0); // visual_mappoint.cc:90-92
        a.accepted[qi] = ok ? 1 : 0;
        if (a.pw) {
            double w[3] = {0, 0, 0};
            if (ok) { // camera.cc:149-151: pose.R * pc + pose.t
                const double pc[3] = {(double) ox, (double) oy, (double) oz};
                double r[3];
                m3_mul_v(a.poseR, pc, r);
                w[0] = r[0] + a.poset[0];
                w[1] = r[1] + a.poset[1];
                w[2] = r[2] + a.poset[2];
            }
            a.pw[3 * qi] = w[0];
            a.pw[3 * qi + 1] = w[1];
            a.pw[3 * qi + 2] = w[2];
        }
    }
    // the exchange step fused into the producer: one 16-byte and one 1-byte store per rank over NVLink (peer memory), no
    // separate collective; k_peer_signal publishes completion afterwards
    if (a.peer.n_ranks) {
#pragma unroll 1
        for (int r = 0; r < a.peer.n_ranks; r++) {
            unsigned char* slot = a.peer.base[r] + a.peer.parity_off + (unsigned long long) a.peer.rank * a.peer.slot_bytes;
            reinterpret_cast<float4*>(slot)[qi] = make_float4(ox, oy, oz, od);
            reinterpret_cast<int8_t*>(slot + 16ull * a.peer.nq_cap)[qi] = (int8_t) type;
        }
    }
}

// PointCloudCommon::pointCloudProjection (pointcloud.cc:50-63): reference operation order, no contraction
__global__ void __launch_bounds__(256) k_project(const PointXYZI* __restrict__ src, PointXYZI* __restrict__ dst, unsigned n,
                                                 const __grid_constant__ Extrinsics e /* Rcl,tcl = the pose */) {
    const unsigned k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const float4* s4 = reinterpret_cast<const float4*>(src + k);
    const float4 p = __ldg(s4), q = __ldg(s4 + 1);
    const double v[3] = {(double) p.x, (double) p.y, (double) p.z};
    double r[3];
    m3_mul_v(e.Rcl, v, r);
    float4* d4 = reinterpret_cast<float4*>(dst + k);
    d4[0] = make_float4((float) (r[0] + e.tcl[0]), (float) (r[1] + e.tcl[1]), (float) (r[2] + e.tcl[2]), p.w);
    d4[1] = q;
}


// ------------------------------------------------------------------------------------------------
// Batched forms of the plane geometry of lidar::PointCloudCommon and of the consumer of the depths in the
// optimiser (rank 4).  One thread per group / factor; reference operation order, no contraction.
// ------------------------------------------------------------------------------------------------
// PointCloudCommon::planeEstimation (pointcloud.cc:73-105): groups of 5 points
__global__ void __launch_bounds__(128) k_plane_estimation(const PointXYZI* __restrict__ pts, unsigned n_groups, double threshold,
                                                          unsigned char* valid, double* unit_normal, double* norm_inverse,
                                                          double* distance, int* n_distance) {
    const unsigned g = blockIdx.x * 128 + threadIdx.x;
    if (g >= n_groups) return;
    double P[5][3];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(pts + 5 * (size_t) g + k));
        P[k][0] = (double) p.x;
        P[k][1] = (double) p.y;
        P[k][2] = (double) p.z;
    }
    double nv[3];
    colpiv_qr_solve_5x3(P, nv);                       // :89
    const double ninv = 1.0 / sqrt(v3_dot(nv, nv));    // :92
    const double un[3] = {nv[0] * ninv, nv[1] * ninv, nv[2] * ninv}; // :93
    int nd = 0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 5; k++) { // :96-103: the distances up to and including the first failing point
        if (ok) {
            const double dis = v3_dot(un, P[k]) + ninv;
            distance[5 * (size_t) g + k] = dis;
            nd++;
            if (fabs(dis) > threshold) ok = false;
        } else {
            distance[5 * (size_t) g + k] = 0.0;
        }
    }
    valid[g] = ok ? 1 : 0;
    unit_normal[3 * (size_t) g] = un[0];
    unit_normal[3 * (size_t) g + 1] = un[1];
    unit_normal[3 * (size_t) g + 2] = un[2];
    norm_inverse[g] = ninv;
    n_distance[g] = nd;
}

// PointCloudCommon::pointToPlaneStd (pointcloud.cc:111-190): groups of `gs` neighbours + one target each
__global__ void __launch_bounds__(128) k_point_to_plane_std(const PointXYZI* __restrict__ nearest, unsigned gs, const PointXYZI* __restrict__ targets,
                                                            unsigned n_groups, double point_std, double* out) {
    const unsigned g = blockIdx.x * 128 + threadIdx.x;
    if (g >= n_groups) return;
    const PointXYZI* __restrict__ pts = nearest + (size_t) g * gs;
    const double pvv = point_std * point_std; // :113-114
    const double inv = 1.0 / (double) gs;     // :117
    double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, center[3] = {0, 0, 0};
    for (unsigned i = 0; i < gs; i++) { // :120-126
        const float4 p = __ldg(reinterpret_cast<const float4*>(pts + i));
        const double pv[3] = {(double) p.x, (double) p.y, (double) p.z};
#pragma unroll
        for (int r = 0; r < 3; r++) {
            center[r] += pv[r];
#pragma unroll
            for (int c = 0; c < 3; c++) cov[3 * r + c] += pv[r] * pv[c];
        }
    }
#pragma unroll
    for (int r = 0; r < 3; r++) center[r] *= inv; // :127
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) cov[3 * r + c] = cov[3 * r + c] * inv - center[r] * center[c]; // :128
    double evals[3], evecs[9];
    sym_eigen3(cov, evals, evecs); // :131-136
    int emin = 0;                  // :138-140
    for (int m = 1; m < 3; m++)
        if (evals[m] < evals[emin]) emin = m;
    double nrm[3], ev[3][3]; // ev[m] = eigenvector m
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
        for (int r = 0; r < 3; r++) ev[m][r] = evecs[3 * r + m];
#pragma unroll
    for (int r = 0; r < 3; r++) nrm[r] = emin == 0 ? ev[0][r] : (emin == 1 ? ev[1][r] : ev[2][r]); // :151
    const double ev_min = emin == 0 ? evals[0] : (emin == 1 ? evals[1] : evals[2]);
    double plane_cov[6][6];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) plane_cov[r][c] = 0;
    for (unsigned i = 0; i < gs; i++) { // :160-176
        const float4 p = __ldg(reinterpret_cast<const float4*>(pts + i));
        const double d[3] = {(double) p.x - center[0], (double) p.y - center[1], (double) p.z - center[2]};
        double F[3][3];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const bool skip = m == emin;
            const double scale = skip ? 0.0 : inv / (ev_min - evals[m]);
            const double de = d[0] * ev[m][0] + d[1] * ev[m][1] + d[2] * ev[m][2], dn = d[0] * nrm[0] + d[1] * nrm[1] + d[2] * nrm[2];
#pragma unroll
            for (int c = 0; c < 3; c++) F[m][c] = skip ? 0.0 : scale * (de * nrm[c] + dn * ev[m][c]);
        }
        double J[6][3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                double sacc = 0;
#pragma unroll
                for (int m = 0; m < 3; m++) sacc += evecs[3 * r + m] * F[m][c];
                J[r][c] = sacc;
            }
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) J[3 + r][c] = r == c ? inv : 0.0;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double sum = 0;
#pragma unroll
                for (int k = 0; k < 3; k++) sum += J[r][k] * J[c][k];
                plane_cov[r][c] += pvv * sum;
            }
    }
    const float4 t = __ldg(reinterpret_cast<const float4*>(targets + g));
    const double jac[6] = {(double) t.x - center[0], (double) t.y - center[1], (double) t.z - center[2], -nrm[0], -nrm[1], -nrm[2]}; // :179-181
    double var = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) var += jac[r] * plane_cov[r][c] * jac[c]; // :184
    var += pvv * (nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);     // :186
    out[g] = sqrt(var);
}

// DepthFactor (factors/depth_factor.h:38-42) as the optimiser adds it (optimizer.cc:659-665: sigma = ASSOCIATE_DEPTH_STD,
// ceres::HuberLoss(1.0)): residual r = (depth - 1 / invdepth) / sigma, its Jacobian d r / d invdepth = 1 / (sigma invdepth^2)
// (what ceres' autodiff yields), and the robustified cost 0.5 rho(r^2) with ceres' Huber rho(s) = s for s <= a^2, else
// 2 a sqrt(s) - a^2 (Ceres Solver >= 2.1.0, README.md:48; loss_function.h).  huber_a <= 0: no loss function.
__global__ void __launch_bounds__(256) k_depth_factor(const double* __restrict__ depth, const double* __restrict__ invdepth, unsigned n,
                                                      double sigma, double huber_a, double* residual, double* jacobian, double* cost) {
    const unsigned k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const double rho_inv = invdepth[k];
    const double r = (depth[k] - 1.0 / rho_inv) / sigma;
    const double s = r * r;
    double rho = s;
    if (huber_a > 0) {
        const double b = huber_a * huber_a;
        if (s > b) rho = 2.0 * huber_a * sqrt(s) - b;
    }
    if (residual) residual[k] = r;
    if (jacobian) jacobian[k] = (1.0 / (rho_inv * rho_inv)) / sigma;
    if (cost) cost[k] = 0.5 * rho;
}

// Optimizer::visualOutlierCullingByChi2 (optimizer.cc:752-783): cost = r^T r per residual block, outlier iff cost > threshold
__global__ void __launch_bounds__(256) k_chi2(const double* __restrict__ residuals, unsigned n, unsigned dim, double threshold, double* cost,
                                              unsigned char* outlier) {
    const unsigned k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    double c = 0;
    for (unsigned i = 0; i < dim; i++) {
        const double r = residuals[(size_t) k * dim + i];
        c += r * r;
    }
    if (cost) cost[k] = c;
    outlier[k] = c > threshold ? 1 : 0;
}

} // namespace lv
