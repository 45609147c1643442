// kernels.cuh -- the four kernels of the LiDAR depth-enhancement path (sm_100a, hand-written, no libraries).
//
//   k_prepare      per keyframe: window search (le_vins.cc:413-437), keyframe pose (le_vins.cc:441-442) and the
//                  per-INS-interval table that replaces the per-point pose interpolation (misc.cc:62-103)
//   k_deskew_bin   per point: range gate + deskew (le_vins.cc:445-464) fused with range-image binning
//                  (tracking.cc:307-336) through a 64-bit atomicMin key (range bits | point index)
//   k_finalize     per cell: winner -> unit vector + range (tracking.cc:340-354), z-buffer reset
//   k_associate    per feature (one warp): exact windowed 5-NN + QR plane fit + validity + depth (tracking.cc:564-634)
// plus k_bin_cloud / k_compact / k_project / k_ring_append for the stage-wise ABI entry points.
#pragma once

#include "device_math.cuh"

namespace lv {

constexpr int DESKEW_THREADS = 256;
constexpr int FINALIZE_THREADS = 256;
constexpr int ASSOC_WARPS = 4;
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
__global__ void k_ring_append(RingState* ring, const PointXYZIT* __restrict__ pts, unsigned long long cap,
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
// k_prepare: one block per keyframe
// ------------------------------------------------------------------------------------------------
struct PrepareArgs {
    KfState* kfs;
    const PointXYZIT* pts;
    unsigned long long pts_cap; // ring capacity (ULLONG_MAX for linear pools)
    const InsSample* ins;
    Interval* table;
    RingState* ring; // streaming: the keyframe's buffer is the ring (le_vins.cc:410-437), and is popped
    Extrinsics ext;
    Consts c;
};

__global__ void __launch_bounds__(128) k_prepare(const __grid_constant__ PrepareArgs a) {
    KfState* kf = a.kfs + blockIdx.x;
    __shared__ double sR0[9], st0[3];
    __shared__ unsigned long long s_first, s_end, s_begin, s_count;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        if (a.ring) {
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

    if (threadIdx.x == 0) {
        const unsigned long long first = s_first;
        const unsigned long long end = s_end > first ? s_end : first;
        kf->proj_begin = wrap(begin + first, a.pts_cap); // the projected slice starts at the first kept point
        kf->first = (unsigned) first;
        kf->n_proj = (unsigned) (end - first);
        kf->status = (end - first) == 0 ? 1 : 0; // le_vins.cc:426-429 / 469
        kf->m = 0;
        kf->ins_t0 = w[0].time;
        kf->ins_t_last = w[nins - 1].time;
        kf->inv_mean_dt = nins > 1 ? (double) (nins - 1) / (w[nins - 1].time - w[0].time) : 0.0;
        if (a.ring) { // the pops are persistent (the deque is shared state)
            a.ring->head = wrap(begin + first, a.pts_cap);
            a.ring->count = count - first;
        }
    }

    // per-interval table
    double R0[9], t0[3];
#pragma unroll
    for (int i = 0; i < 9; i++) R0[i] = sR0[i];
    t0[0] = st0[0];
    t0[1] = st0[1];
    t0[2] = st0[2];
    Interval* __restrict__ tab = a.table + kf->table_begin;
    for (int j = threadIdx.x; j < nins; j += blockDim.x) {
        Interval e;
        double Rq[9], C[9], d[3];
        if (j == 0) { // out-of-window fallback: newest state, no interpolation (misc.cc:74-77)
            const InsSample s = w[nins - 1];
            quat_to_matrix(s.q, Rq);
            e.t_lo = e.t_hi = 0;
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
        e.pad = 0;
        tab[j] = e;
    }
}

// ------------------------------------------------------------------------------------------------
// k_deskew_bin: one thread per point
// ------------------------------------------------------------------------------------------------
struct DeskewArgs {
    const KfState* kfs; // first keyframe of this launch (blockIdx.y indexes from here)
    const PointXYZIT* pts;
    unsigned long long pts_cap;
    const Interval* table;
    unsigned long long* zbuf; // [gridDim.y][CELLS]
    PointXYZI* cloud_out;     // streaming only (gridDim.y == 1), nullable
    Extrinsics ext;
    Consts c;
};

__device__ __forceinline__ void zbuf_min(unsigned long long* cell, unsigned long long key) {
    // most points lose: filter with a plain L2 read first (a stale value is only ever larger, so no update is missed)
    if (key < __ldcg(cell)) atomicMin(cell, key);
}

template <int LIBM_MODE, bool WRITE_CLOUD>
__global__ void __launch_bounds__(DESKEW_THREADS) k_deskew_bin(const __grid_constant__ DeskewArgs a) {
    const KfState* __restrict__ kf = a.kfs + blockIdx.y;
    const unsigned n = kf->n_proj;
    const unsigned k = blockIdx.x * DESKEW_THREADS + threadIdx.x;
    if (blockIdx.x * DESKEW_THREADS >= n) return;
    if (k >= n) return;
    KfView kv;
    kv.tab = a.table + kf->table_begin;
    kv.n_ins = (int) kf->ins_count;
    kv.t0 = kf->ins_t0;
    kv.t_last = kf->ins_t_last;
    kv.inv_mean_dt = kf->inv_mean_dt;

    const unsigned long long pi = wrap(kf->proj_begin + k, a.pts_cap);
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(a.pts + pi);
    const float4 xyz = __ldg(p4);
    const double2 it = __ldg(reinterpret_cast<const double2*>(p4 + 1));

    float ox, oy, oz;
    const bool kept = deskew_point(xyz.x, xyz.y, xyz.z, it.y, kv, a.ext, a.c, ox, oy, oz);
    if (WRITE_CLOUD) {
        float4* out = reinterpret_cast<float4*>(a.cloud_out + k);
        out[0] = kept ? make_float4(ox, oy, oz, 1.0f) : make_float4(0.f, 0.f, 0.f, 1.0f); // resize() default (le_vins.cc:438)
        out[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!kept) return; // a default point lands at col = -150 and is dropped at tracking.cc:326-328
    const int cell = cell_of_point<LIBM_MODE>(ox, oy, oz, a.c);
    if (cell < 0) return;
    const float range = sqrtf(v3f_sqnorm(ox, oy, oz)); // tracking.cc:330
    if (!(range < FLT_MAX)) return;                    // tracking.cc:332 against the FLT_MAX initial value
    const unsigned long long key = ((unsigned long long) __float_as_uint(range) << 32) | k;
    zbuf_min(a.zbuf + (size_t) blockIdx.y * CELLS + cell, key);
}

// explicit-cloud binning for levins_build_range_image(cloud != NULL)
template <int LIBM_MODE>
__global__ void __launch_bounds__(256) k_bin_cloud(const PointXYZI* __restrict__ cloud, unsigned n, unsigned long long* zbuf,
                                                   const __grid_constant__ Consts c) {
    const unsigned k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const float4 p = __ldg(reinterpret_cast<const float4*>(cloud + k));
    const int cell = cell_of_point<LIBM_MODE>(p.x, p.y, p.z, c);
    if (cell < 0) return;
    const float range = sqrtf(v3f_sqnorm(p.x, p.y, p.z));
    if (!(range < FLT_MAX)) return;
    zbuf_min(zbuf + cell, ((unsigned long long) __float_as_uint(range) << 32) | k);
}

// ------------------------------------------------------------------------------------------------
// k_finalize: one thread per cell.  cells[] = (unit x, unit y, unit z, range), range = -1 marks an empty cell.
// ------------------------------------------------------------------------------------------------
struct FinalizeArgs {
    KfState* kfs;
    const PointXYZIT* pts;
    unsigned long long pts_cap;
    const Interval* table;
    const PointXYZI* cloud; // FROM_CLOUD source
    unsigned long long* zbuf;
    float4* cells;   // [gridDim.y][CELLS]
    int32_t* winner; // nullable, [CELLS] (single keyframe)
    Extrinsics ext;
    Consts c;
};

template <bool FROM_CLOUD>
__global__ void __launch_bounds__(FINALIZE_THREADS) k_finalize(const __grid_constant__ FinalizeArgs a) {
    KfState* kf = a.kfs + blockIdx.y;
    const int cell = blockIdx.x * FINALIZE_THREADS + threadIdx.x;
    bool filled = false;
    if (cell < CELLS) {
        unsigned long long* zp = a.zbuf + (size_t) blockIdx.y * CELLS + cell;
        const unsigned long long key = *zp;
        float4 out = make_float4(0.f, 0.f, 0.f, -1.0f);
        int win = -1;
        if (key != EMPTY_KEY) {
            const unsigned k = (unsigned) (key & 0xffffffffu);
            float x, y, z;
            if (FROM_CLOUD) {
                const float4 p = __ldg(reinterpret_cast<const float4*>(a.cloud + k));
                x = p.x;
                y = p.y;
                z = p.z;
            } else { // recompute the winner's deskew: same function, same bits as in k_deskew_bin
                KfView kv;
                kv.tab = a.table + kf->table_begin;
                kv.n_ins = (int) kf->ins_count;
                kv.t0 = kf->ins_t0;
                kv.t_last = kf->ins_t_last;
                kv.inv_mean_dt = kf->inv_mean_dt;
                const unsigned long long pi = wrap(kf->proj_begin + k, a.pts_cap);
                const float4* __restrict__ p4 = reinterpret_cast<const float4*>(a.pts + pi);
                const float4 xyz = __ldg(p4);
                const double2 it = __ldg(reinterpret_cast<const double2*>(p4 + 1));
                deskew_point(xyz.x, xyz.y, xyz.z, it.y, kv, a.ext, a.c, x, y, z);
            }
            const float n2 = v3f_sqnorm(x, y, z); // tracking.cc:348-349
            const float r = sqrtf(n2);
            if (n2 > 0.f) {
                const float nn = sqrtf(n2);
                out = make_float4(x / nn, y / nn, z / nn, r);
            } else {
                out = make_float4(x, y, z, r);
            }
            win = (int) k;
            filled = true;
            *zp = EMPTY_KEY; // leave the z-buffer ready for the next keyframe that uses this slot
        }
        a.cells[(size_t) blockIdx.y * CELLS + cell] = out;
        if (a.winner) a.winner[cell] = win;
    }
    const unsigned cnt = __syncthreads_count(filled);
    if (threadIdx.x == 0 && cnt) atomicAdd(&kf->m, cnt);
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
// k_associate: one warp per feature
// ------------------------------------------------------------------------------------------------
struct AssocArgs {
    const KfState* kfs;
    const float* queries; // pool, 3 floats per query
    const float4* cells;  // [gridDim.y][CELLS]
    const int32_t* c2c;   // nullable (single keyframe): cell -> compact index
    int8_t* type;
    float4* xyz_depth;
    int32_t* flags;
    int32_t* knn_idx; // nullable, 5 per query
    float* knn_d2;    // nullable
    Consts c;
};

__global__ void __launch_bounds__(ASSOC_WARPS * 32) k_associate(const __grid_constant__ AssocArgs a) {
    const KfState* __restrict__ kf = a.kfs + blockIdx.y;
    const int lane = threadIdx.x & 31;
    const unsigned f = blockIdx.x * ASSOC_WARPS + (threadIdx.x >> 5);
    if (f >= kf->q_count) return;
    const size_t qi = kf->q_begin + f;
    const float qx = __ldg(a.queries + 3 * qi), qy = __ldg(a.queries + 3 * qi + 1), qz = __ldg(a.queries + 3 * qi + 2);
    const float4* __restrict__ cells = a.cells + (size_t) blockIdx.y * CELLS;

    // the query's own cell; precision is irrelevant here, the window has a full cell of slack
    const float el = atan2f(qy, sqrtf(qx * qx + qz * qz));
    const float az = atan2f(qz, qx);
    const float inv_ang = (float) a.c.inv_angle, off = (float) a.c.half_offset;
    const int rowq = (int) rintf((el + off) * inv_ang);
    const int colq = (int) rintf((az - off) * inv_ang);

    unsigned dbits[4];
    int cid[4];
#pragma unroll
    for (int itc = 0; itc < 4; itc++) {
        dbits[itc] = 0x7f800000u; // +inf
        cid[itc] = 0x7fffffff;
        const int j = lane + 32 * itc;
        if (j < WIN_CELLS) {
            const int r = rowq - WIN_ROWS + j / WIN_W;
            const int cc = colq - WIN_COLS + j % WIN_W;
            if (r >= 0 && r < IMG && cc >= 0 && cc < IMG) {
                const float4 v = __ldg(cells + r * IMG + cc);
                if (v.w >= 0.f) {
                    // FLANN L2_Simple<float>: result = 0; result += diff*diff, x then y then z
                    const float dx = qx - v.x, dy = qy - v.y, dz = qz - v.z;
                    float d = dx * dx;
                    d = d + dy * dy;
                    d = d + dz * dz;
                    dbits[itc] = __float_as_uint(d);
                    cid[itc] = r * IMG + cc;
                }
            }
        }
    }

    // top-5 by (d2, cell id): cell ids are row-major, i.e. the order of the reference's compacted indices
    unsigned nn_d[5];
    int nn_c[5];
    int found = 0;
#pragma unroll
    for (int r = 0; r < 5; r++) {
        unsigned bd = dbits[0];
        int bc = cid[0];
#pragma unroll
        for (int itc = 1; itc < 4; itc++)
            if (dbits[itc] < bd || (dbits[itc] == bd && cid[itc] < bc)) {
                bd = dbits[itc];
                bc = cid[itc];
            }
        const unsigned m1 = __reduce_min_sync(0xffffffffu, bd);
        const unsigned m2 = __reduce_min_sync(0xffffffffu, bd == m1 ? (unsigned) bc : 0xffffffffu);
        nn_d[r] = m1;
        nn_c[r] = (int) m2;
        if (m1 != 0x7f800000u) found++;
        if (bd == m1 && (unsigned) bc == m2) { // this lane owns the winner: retire it
#pragma unroll
            for (int itc = 0; itc < 4; itc++)
                if (dbits[itc] == m1 && cid[itc] == (int) m2) {
                    dbits[itc] = 0x7f800000u;
                    cid[itc] = 0x7fffffff;
                }
        }
    }

    int type = -1, flags = 0;
    float ox = qx, oy = qy, oz = qz, od = 0.f;
    if (found < 5) {
        flags = 1;
    } else if ((double) __uint_as_float(nn_d[4]) > a.c.thr2) { // tracking.cc:573-575
        flags = 2;
    } else {
        double P[5][3];
        float rng0 = 0.f, rng4 = 0.f;
#pragma unroll
        for (int k = 0; k < 5; k++) { // tracking.cc:577-581
            const float4 v = __ldg(cells + nn_c[k]);
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
    if (lane == 0) {
        a.type[qi] = (int8_t) type;
        a.xyz_depth[qi] = make_float4(ox, oy, oz, od);
        a.flags[qi] = flags;
    }
    if (a.knn_idx && lane < 5) {
        // nn_* are warp-uniform; select without dynamic register indexing
        unsigned d = nn_d[0];
        int cI = nn_c[0];
#pragma unroll
        for (int k = 1; k < 5; k++)
            if (lane == k) {
                d = nn_d[k];
                cI = nn_c[k];
            }
        const bool ok = d != 0x7f800000u;
        a.knn_idx[5 * qi + lane] = ok ? (a.c2c ? a.c2c[cI] : cI) : -1;
        if (a.knn_d2) a.knn_d2[5 * qi + lane] = ok ? __uint_as_float(d) : FLT_MAX;
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

__global__ void k_fill_u64(unsigned long long* p, size_t n, unsigned long long v) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

} // namespace lv
