// device_math.cuh -- device-side arithmetic of the LiDAR depth-enhancement path (sm_100a).
//
// This translation unit is compiled with -fmad=false: wherever the reference's result is an integer / index /
// validity decision (range gate, cell id, k-NN order, plane validity, depth in f32) the operation order of the
// reference's x86-64 SSE2 build is reproduced exactly and must not be contracted.  Where only a tolerance applies
// (stage 1, FP64 rigid-body algebra) fused multiply-adds are requested explicitly with fma().
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace lv {

constexpr int IMG = 300;
constexpr int CELLS = IMG * IMG;
constexpr unsigned long long EMPTY_KEY = ~0ull;

// ---- records (layout-identical to include/levins/abi.h) ----
struct __align__(16) PointXYZIT {
    float x, y, z, pad0;
    float intensity, pad1;
    double time;
};
struct __align__(16) PointXYZI {
    float x, y, z, w;
    float intensity, pad[3];
};
struct InsSample {
    double time, state_time;
    double p[3];
    double q[4];
};
static_assert(sizeof(PointXYZIT) == 32 && sizeof(PointXYZI) == 32 && sizeof(InsSample) == 72, "ABI layout");

// One INS interval [i-1, i] pre-digested for the per-point deskew ("stage 1 algebra"):
//   p_c0 = A * Rod(k, -s*theta) * (R_bc p_c1 + t_bc) + c0 + s * c1,   s = (t - ts0) * inv_dts
// entry 0 of each keyframe's table is the out-of-window fallback (newest state, misc.cc:74-77); its (t_lo, t_hi) are
// (time of the last sample, time of the first sample), which makes entry_brackets() true exactly for the times OUTSIDE
// the window (see entry_brackets).
struct __align__(16) Interval {
    double t_lo, t_hi;   // IMU times of samples i-1, i (search keys, misc.cc:40-43)
    double ts0, inv_dts; // state time of sample i-1 and 1 / (ts1 - ts0)   (misc.cc:88)
    double A[9];         // R0^T R(q0), row-major
    double k[3];         // rotation axis of dq = q1^-1 q0   (misc.cc:85-86)
    double theta;        // its angle
    double c0[3];        // R0^T (p0 - t0)
    double c1[3];        // R0^T (p1 - p0)
    double fast;         // != 0: entry_brackets() may be used on (t_lo, t_hi) and theta is in sin_omc_small()'s range
    // Certified frustum cull (k_deskew_bin): the whole deskew of this interval frozen at mid-interval, in f32:
    //   p~ = M p_lidar + m,   |p~ - p_c0| <= ca |p_lidar| + cb   for every time the entry brackets
    // (rotation and translation left over within the interval + every rounding, all rounded up).  A point whose p~ lies
    // outside the 90-degree image cone by more than that bound is rejected by tracking.cc:311-314 whatever its exact
    // coordinates are, so the f64 algebra is skipped for it.  ca < 0 marks "no cull" (bound not finite).
    float M[9]; // row-major
    float m[3];
    float ca, cb;
    float pad[2];
};
static_assert(sizeof(Interval) == 256, "interval entry");

// fixed extrinsics (class VINS: pose_c_l_, pose_b_c_, td_b_l_), R column-major like Eigen
struct Extrinsics {
    double Rcl[9], tcl[3];
    double Rbc[9], tbc[3];
    double td_b_l;
};

// constants evaluated on the host with the reference's own double expressions (tracking.h:119-135)
struct Consts {
    double angle;       // ASSOCIATE_RANGE_IMAGE_ANGLE
    double inv_angle;   // 1 / angle (fast path of the cell computation; exactness restored by a guarded fallback)
    double half_offset; // M_PI * 0.5 - ASSOCIATE_IMAGE_FOV * 0.5
    double tan_value;   // ASSOCIATE_POINT_TAN_VALUE
    double thr2;        // ASSOCIATE_NEAREST_SQUARE_DISTANCE_THRESHOLD
    double nearest, farthest;
    double plane_residual, range_spread;
    double stack_interval;
    int libm_mode;
    int pad;
    // exact f32 restatements of the range gate on the squared norm (sqrtf is monotone and correctly rounded):
    //   (double) sqrtf(n2) > farthest  <=>  n2 > n2_far        (double) sqrtf(n2) < nearest  <=>  n2 < n2_near
    float n2_far, n2_near;
    float inv_angle_f; // (float) (1 / ANGLE)
    float cell_guard;  // fast-path cell ids closer than this (in cells) to a rounding boundary are recomputed exactly
};

// per-keyframe state, device resident (a whole number of 16-byte chunks: k_deskew_bin prefetches it with cp.async)
struct __align__(16) KfState {
    unsigned long long pt_begin, pt_count;
    unsigned long long ins_begin, ins_count;
    unsigned long long q_begin, q_count;
    double stamp;
    unsigned long long table_begin; // this keyframe's slice of the interval-table pool (ins_count entries)
    // written by k_prepare
    unsigned long long proj_begin; // physical index of the first kept point (pt_begin + first, wrapped)
    unsigned int first, n_proj;
    int status; // LEVINS_OK / LEVINS_EMPTY
    unsigned int m;
    double ins_t0, ins_t_last, inv_mean_dt;
    unsigned int ticket; // k_prepare blocks of this keyframe that have finished (streaming: the last one pops the ring)
    unsigned int pad[3];
};
static_assert(sizeof(KfState) % 16 == 0 && sizeof(KfState) <= 128, "KfState is copied in 16-byte chunks");

// ------------------------------------------------------------------------------------------------
// small algebra (column-major 3x3 like Eigen), mirrored summation order ((a0 b0 + a1 b1) + a2 b2)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void m3_mul_v(const double* R, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 3; i++) y[i] = (R[i] * x[0] + R[3 + i] * x[1]) + R[6 + i] * x[2];
}
__host__ __device__ __forceinline__ void m3t_mul_v(const double* R, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 3; i++) y[i] = (R[3 * i] * x[0] + R[3 * i + 1] * x[1]) + R[3 * i + 2] * x[2];
}
__host__ __device__ __forceinline__ void m3_mul_m(const double* A, const double* B, double* C) {
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) C[3 * j + i] = (A[i] * B[3 * j] + A[3 + i] * B[3 * j + 1]) + A[6 + i] * B[3 * j + 2];
}
__host__ __device__ __forceinline__ void m3t_mul_m(const double* A, const double* B, double* C) {
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            C[3 * j + i] = (A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1]) + A[3 * i + 2] * B[3 * j + 2];
}
__host__ __device__ __forceinline__ double v3_dot(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

__host__ __device__ __forceinline__ void quat_mul(const double* a, const double* b, double* r) {
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
    r[3] = aw * bw - ax * bx - ay * by - az * bz;
    r[0] = aw * bx + ax * bw + ay * bz - az * by;
    r[1] = aw * by + ay * bw + az * bx - ax * bz;
    r[2] = aw * bz + az * bw + ax * by - ay * bx;
}
__host__ __device__ __forceinline__ double quat_sqnorm(const double* q) { return ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]; }
__host__ __device__ __forceinline__ void quat_inverse(const double* q, double* r) {
    const double n2 = quat_sqnorm(q);
    if (n2 > 0) {
        r[0] = -q[0] / n2;
        r[1] = -q[1] / n2;
        r[2] = -q[2] / n2;
        r[3] = q[3] / n2;
    } else {
        r[0] = r[1] = r[2] = r[3] = 0;
    }
}
__host__ __device__ __forceinline__ void quat_normalize(double* q) {
    const double z = quat_sqnorm(q);
    if (z > 0) {
        const double s = sqrt(z);
        q[0] /= s;
        q[1] /= s;
        q[2] /= s;
        q[3] /= s;
    }
}
// Eigen toRotationMatrix, column-major output
__host__ __device__ __forceinline__ void quat_to_matrix(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz);
    R[3] = txy - twz;
    R[6] = txz + twy;
    R[1] = txy + twz;
    R[4] = 1 - (txx + tzz);
    R[7] = tyz - twx;
    R[2] = txz - twy;
    R[5] = tyz + twx;
    R[8] = 1 - (txx + tyy);
}

// Eigen::AngleAxisd(quaternion) (rotation.cc:65-68): angle and unit axis
__host__ __device__ __forceinline__ void quat_to_angle_axis(const double* q, double& angle, double* axis) {
    double n = sqrt(v3_dot(q, q));
    if (n < DBL_EPSILON) {
        const double m = fmax(fabs(q[0]), fmax(fabs(q[1]), fabs(q[2])));
        if (m > 0) {
            const double a = q[0] / m, b = q[1] / m, c = q[2] / m;
            n = m * sqrt((a * a + b * b) + c * c);
        } else {
            n = 0;
        }
    }
    if (n != 0) {
        angle = 2 * atan2(n, fabs(q[3]));
        if (q[3] < 0) n = -n;
        axis[0] = q[0] / n;
        axis[1] = q[1] / n;
        axis[2] = q[2] / n;
    } else {
        angle = 0;
        axis[0] = 1;
        axis[1] = 0;
        axis[2] = 0;
    }
}

// rotation.cc:59-63
__host__ __device__ __forceinline__ void rotvec_to_quat(const double* rvec, double* q) {
    const double z = v3_dot(rvec, rvec);
    const double angle = sqrt(z);
    double vec[3] = {rvec[0], rvec[1], rvec[2]};
    if (z > 0) {
        const double s = sqrt(z);
        vec[0] /= s;
        vec[1] /= s;
        vec[2] /= s;
    }
    const double ha = 0.5 * angle;
    double sn, cs;
    sincos(ha, &sn, &cs);
    q[3] = cs;
    q[0] = sn * vec[0];
    q[1] = sn * vec[1];
    q[2] = sn * vec[2];
}

// misc.cc:80-95: interpolated state (p, q) between two samples, reference operation order
__host__ __device__ inline void interpolate_state(const InsSample& s0, const InsSample& s1, double midtime, double* p, double* q) {
    double dp[3] = {s1.p[0] - s0.p[0], s1.p[1] - s0.p[1], s1.p[2] - s0.p[2]};
    double q1inv[4], dq[4], rvec[3], axis[3], angle;
    quat_inverse(s1.q, q1inv);
    quat_mul(q1inv, s0.q, dq);
    quat_to_angle_axis(dq, angle, axis);
    rvec[0] = angle * axis[0];
    rvec[1] = angle * axis[1];
    rvec[2] = angle * axis[2];
    const double scale = (midtime - s0.state_time) / (s1.state_time - s0.state_time);
    rvec[0] = rvec[0] * scale;
    rvec[1] = rvec[1] * scale;
    rvec[2] = rvec[2] * scale;
    rotvec_to_quat(rvec, dq);
    p[0] = s0.p[0] + dp[0] * scale;
    p[1] = s0.p[1] + dp[1] * scale;
    p[2] = s0.p[2] + dp[2] * scale;
    double dqinv[4];
    quat_inverse(dq, dqinv);
    quat_mul(s0.q, dqinv, q);
    quat_normalize(q);
}
// misc.cc:97-103: pose of the sensor `ext` rides on the body state (p, q)
__host__ __device__ inline void pose_with_extrinsic(const double* p, const double* q, const double* Rext, const double* text, double* R,
                                                    double* t) {
    double Rq[9], Rt[3];
    quat_to_matrix(q, Rq);
    m3_mul_v(Rq, text, Rt);
    t[0] = p[0] + Rt[0];
    t[1] = p[1] + Rt[1];
    t[2] = p[2] + Rt[2];
    m3_mul_m(Rq, Rext, R);
}
// misc.cc:80-95 then 97-103 (used once per keyframe for the pose at `stamp`)
__host__ __device__ inline void interpolate_pose(const InsSample& s0, const InsSample& s1, double midtime, const double* Rbc,
                                        const double* tbc, double* R, double* t) {
    double p[3], q[4];
    interpolate_state(s0, s1, midtime, p, q);
    pose_with_extrinsic(p, q, Rbc, tbc, R, t);
}
__host__ __device__ inline void state_pose(const InsSample& s, const double* Rbc, const double* tbc, double* R, double* t) {
    pose_with_extrinsic(s.p, s.q, Rbc, tbc, R, t);
}

// ------------------------------------------------------------------------------------------------
// fdlibm-lineage atanf / atan2f: the float overloads the reference's unqualified atan2 calls resolve to on
// glibc (tracking.cc:317-320): glibc's sysdeps/ieee754/flt-32 algorithm, operation for operation; no contraction
// (-fmad=false), IEEE division.  The CPU checker's copy of the same sequence is verified bit-identical to glibc 2.39's
// atan2f on 2e7 inputs in tests/test_oracle.py, and this device copy against the checker in the GPU parity tests.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float atanf_fd(float x) {
    const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
    const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
    const float c0 = 3.3333334327e-01f, c1 = -2.0000000298e-01f, c2 = 1.4285714924e-01f, c3 = -1.1111110449e-01f,
                c4 = 9.0908870101e-02f, c5 = -7.6918758452e-02f, c6 = 6.6610731184e-02f, c7 = -5.8335702866e-02f,
                c8 = 4.9768779427e-02f, c9 = -3.6531571299e-02f, c10 = 1.6285819933e-02f;
    const int hx = __float_as_int(x);
    const int ix = hx & 0x7fffffff;
    float hi = 0.f, lo = 0.f;
    int id;
    if (ix >= 0x4c000000) {
        if (ix > 0x7f800000) return x + x;
        if (hx > 0) return hi3 + lo3;
        return -hi3 - lo3;
    }
    if (ix < 0x3ee00000) {
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) {
                id = 0;
                hi = hi0;
                lo = lo0;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else {
                id = 1;
                hi = hi1;
                lo = lo1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000) {
                id = 2;
                hi = hi2;
                lo = lo2;
                x = (x - 1.5f) / (1.0f + 1.5f * x);
            } else {
                id = 3;
                hi = hi3;
                lo = lo3;
                x = -1.0f / x;
            }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (c0 + w * (c2 + w * (c4 + w * (c6 + w * (c8 + w * c10)))));
    const float s2 = w * (c1 + w * (c3 + w * (c5 + w * (c7 + w * c9))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return (hx < 0) ? -r : r;
}

__device__ __forceinline__ float atan2f_fd(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    const int hx = __float_as_int(x), hy = __float_as_int(y);
    const int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanf_fd(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
        case 0:
        case 1: return y;
        case 2: return pi + tiny;
        default: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
            case 0: return pi_o_4 + tiny;
            case 1: return -pi_o_4 - tiny;
            case 2: return 3.0f * pi_o_4 + tiny;
            default: return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return pi + tiny;
            default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60)
        z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60)
        z = 0.0f;
    else
        z = atanf_fd(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return __int_as_float(__float_as_int(z) ^ (int) 0x80000000);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

// Eigen Vector3f redux order: e0 + (e1 + e2)
// ---- packed f32x2 arithmetic (Blackwell FMUL2 / FADD2 / FFMA2: two IEEE round-to-nearest operations per issue slot;
// each half is bit-identical to the scalar instruction) ----
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; "
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 f2_splat(float v) { return make_float2(v, v); }

// Eigen's Vector3f::squaredNorm() order: x*x + (y*y + z*z), no contraction
__device__ __forceinline__ float v3f_sqnorm(float x, float y, float z) { return x * x + (y * y + z * z); }
// the same bits with y*y and z*z as one packed multiply: measured faster in the single-precision instantiation of
// k_deskew_bin (-15 us), slower in the f64 one (+3 us), so the kernel picks by mode
template <bool PACKED>
__device__ __forceinline__ float v3f_sqnorm_t(float x, float y, float z) {
    if (PACKED) {
        const float2 q = f2_mul(make_float2(y, z), make_float2(y, z));
        return x * x + (q.x + q.y);
    }
    return x * x + (y * y + z * z);
}

// ------------------------------------------------------------------------------------------------
// stage 2: cell of a camera-frame point (tracking.cc:311-331).  Returns -1 when the point is rejected.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int angle_to_index(double a, const Consts& c) {
    // (int) round(a / ANGLE) (tracking.cc:323-324).  Fast path: multiply by the reciprocal; whenever the product is
    // close enough to a rounding boundary for the last-bit difference to matter, redo it with the true division.
    const double qf = a * c.inv_angle;
    const double r = round(qf);
    if (fabs(fabs(qf - r) - 0.5) < 1e-9) return (int) round(a / c.angle);
    return (int) r;
}

template <int LIBM_MODE>
__device__ __forceinline__ int cell_of_point(float x, float y, float z, const Consts& c) {
    if ((z < 0.f) || ((double) fabsf(x / z) > c.tan_value) || ((double) fabsf(y / z) > c.tan_value)) return -1;
    const float s = x * x + z * z;
    double row_angle, col_angle;
    if (LIBM_MODE == 0) {
        row_angle = (double) atan2f_fd(y, sqrtf(s)) + c.half_offset;
        col_angle = (double) atan2f_fd(z, x) - c.half_offset;
    } else {
        row_angle = atan2((double) y, sqrt((double) s)) + c.half_offset;
        col_angle = atan2((double) z, (double) x) - c.half_offset;
    }
    const int row = angle_to_index(row_angle, c);
    const int col = angle_to_index(col_angle, c);
    if ((row < 0) || (col < 0) || (row >= IMG) || (col >= IMG)) return -1;
    return col + row * IMG;
}

// ------------------------------------------------------------------------------------------------
// z-buffer key: [gen:8][range bits:31][point index:25].  atomicMin keeps, per cell, the nearest point and on
// equal range the first index (tracking.cc:332: strictly smaller range replaces).  `gen` counts DOWN from 254 once
// per use of a z-buffer, so keys left by earlier keyframes (larger gen) lose against any new key and read as
// "empty": the buffer never has to be cleared between keyframes (only every 254 uses).
// ------------------------------------------------------------------------------------------------
constexpr unsigned KEY_IDX_BITS = 25;
constexpr unsigned KEY_IDX_MASK = (1u << KEY_IDX_BITS) - 1;
constexpr unsigned long long MAX_POINTS_PER_KEYFRAME = 1ull << KEY_IDX_BITS;
__device__ __forceinline__ unsigned long long key_make(unsigned gen, float range, unsigned idx) {
    return ((unsigned long long) gen << 56) | ((unsigned long long) __float_as_uint(range) << KEY_IDX_BITS) | idx;
}
__device__ __forceinline__ bool key_live(unsigned long long key, unsigned gen) { return (unsigned) (key >> 56) == gen; }
__device__ __forceinline__ unsigned key_index(unsigned long long key) { return (unsigned) key & KEY_IDX_MASK; }

// ------------------------------------------------------------------------------------------------
// stage 2 fast path.  The cell id only needs the angles to ~1e-4 cell except next to a rounding boundary, so the
// common case uses a 9-term odd minimax polynomial for atan on [-1, 1] (|error| < 6e-9 + f32 rounding) on
// t = x/z resp. t = y/sqrt(x^2+z^2) (inside the FOV gate both lie in [-1, 1]), and every id that falls within
// `cell_guard` of a boundary is recomputed with the exact, mirrored cell_of_point<>().  Result: bit-identical cell ids.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float atan_poly(float t) {
    const float u = t * t;
    float p = 0.0024566929787397385f;
    p = fmaf(p, u, -0.014401225373148918f);
    p = fmaf(p, u, 0.03978099301457405f);
    p = fmaf(p, u, -0.07234836369752884f);
    p = fmaf(p, u, 0.10498934984207153f);
    p = fmaf(p, u, -0.14161226153373718f);
    p = fmaf(p, u, 0.19985906779766083f);
    p = fmaf(p, u, -0.33332598209381104f);
    p = fmaf(p, u, 0.9999998807907104f);
    return t * p;
}

// FOV gate of tracking.cc:311-314 without the two IEEE divisions:
//   fl32(|x| / |z|) > TAN (TAN = 1 - 2^-53)  <=>  fl32(|x| / |z|) >= 1  <=>  |x| / |z| >= 1 - 2^-25  (round to nearest even)
//   <=>  |x| - |z| >= -|z| 2^-25, where the left side is exact in f32 by Sterbenz's lemma whenever it matters.
// z == 0 rejects too (x/0 = inf; 0/0 = NaN passes the reference's gate but such a point then fails the col >= 0 test).
__device__ __forceinline__ bool outside_fov(float x, float y, float z) {
    const float az = fabsf(z);
    const float lim = -az * 2.98023223876953125e-8f; // 2^-25
    // (|x| - az >= lim) or (|y| - az >= lim)  <=>  max(|x|, |y|) - az >= lim: the rounded subtraction is monotone
    return (z < 0.f) | (fmaxf(fabsf(x), fabsf(y)) - az >= lim);
}

// atan for the pair (t.x, t.y): atan_poly() on both halves at once
__device__ __forceinline__ float2 atan_poly2(float2 t) {
    const float2 u = f2_mul(t, t);
    float2 p = f2_splat(0.0024566929787397385f);
    p = f2_fma(p, u, f2_splat(-0.014401225373148918f));
    p = f2_fma(p, u, f2_splat(0.03978099301457405f));
    p = f2_fma(p, u, f2_splat(-0.07234836369752884f));
    p = f2_fma(p, u, f2_splat(0.10498934984207153f));
    p = f2_fma(p, u, f2_splat(-0.14161226153373718f));
    p = f2_fma(p, u, f2_splat(0.19985906779766083f));
    p = f2_fma(p, u, f2_splat(-0.33332598209381104f));
    p = f2_fma(p, u, f2_splat(0.9999998807907104f));
    return f2_mul(t, p);
}

// returns the cell id or -1; sets `ambiguous` when the exact path must decide.  Straight-line code (no early exits) so
// that several points per thread interleave.  The column and the row travel as the two halves of packed f32x2
// operations; 1/z and 1/sqrt(x^2+z^2) are single MUFU approximations (this path only has to be right to ~1e-4 cell);
// rounding uses the 1.5 * 2^23 magic constant instead of FRND / F2I.
__device__ __forceinline__ int cell_of_point_fast(float x, float y, float z, const Consts& c, bool& ambiguous) {
    const bool behind = z < 0.f;
    // z outside [2^-61, 2^59) (zero, denormal range, huge, NaN; negative too, but `behind` wins): 1/z or z*z may leave
    // the normal range, so the exact path decides.  Never the case for a point that passed the range gate.
    const bool odd = (__float_as_uint(z) - 0x21000000u) >= 0x3c000000u;
    const bool outside = outside_fov(x, y, z);
    float rz, rs;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rz) : "f"(z));
    const float s = fmaf(x, x, z * z);
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s));
    // col_angle / ANGLE = (atan2(z, x) - pi/4) / ANGLE = 150 - atan(x / z) / ANGLE
    // row_angle / ANGLE = (atan2(y, sqrt(x^2 + z^2)) + pi/4) / ANGLE = 150 + atan(y / sqrt(x^2 + z^2)) / ANGLE
    const float2 at = atan_poly2(f2_mul(make_float2(x, y), make_float2(rz, rs)));
    const float2 q = f2_fma(at, make_float2(-c.inv_angle_f, c.inv_angle_f), f2_splat(150.0f));
    const float MAGIC = 12582912.0f; // 1.5 * 2^23: (q + MAGIC) - MAGIC = rint(q) for |q| < 2^22
    const float2 m = f2_add(q, f2_splat(MAGIC));
    const float2 r = f2_add(m, f2_splat(-MAGIC));
    const float2 d = f2_add(q, make_float2(-r.x, -r.y));
    const int col = __float_as_int(m.x) - 0x4B400000, row = __float_as_int(m.y) - 0x4B400000;
    const bool near_edge = fmaxf(fabsf(d.x), fabsf(d.y)) > 0.5f - c.cell_guard;
    const bool in_grid = ((unsigned) row < (unsigned) IMG) && ((unsigned) col < (unsigned) IMG);
    ambiguous = !behind && (odd || (!outside && near_edge));
    return (behind || odd || outside || near_edge || !in_grid) ? -1 : col + row * IMG;
}

// ------------------------------------------------------------------------------------------------
// stage 1: interval lookup (misc.cc:25-60 semantics) and per-point deskew
// ------------------------------------------------------------------------------------------------
// Returns the table entry index: i in [1, n-1] with time[i-1] <= t < time[i], or 0 (fallback) outside the window.
struct KfView {
    const Interval* tab;
    int n_ins;
    double t0, t_last, inv_mean_dt;
};

template <bool NC = true> // NC: the table is read-only for this kernel (ld.global.nc); false inside the kernel that writes it
__device__ __noinline__ int find_interval_slow(const Interval* __restrict__ tab, int n_ins, double t, int g, int lo, int hi) {
#pragma unroll 1
    for (int it = 0; it < 48; it++) {
        double2 ab;
        if (NC) {
            ab = __ldg(reinterpret_cast<const double2*>(tab + g));
        } else {
            const volatile double* pv = reinterpret_cast<const volatile double*>(tab + g);
            ab.x = pv[0];
            ab.y = pv[1];
        }
        if (t < ab.x) {
            hi = g - 1;
            g = (it < 2) ? g - 1 : (lo + hi) >> 1;
        } else if (t >= ab.y) {
            lo = g + 1;
            g = (it < 2) ? g + 1 : (lo + hi) >> 1;
        } else {
            return g;
        }
        if (lo > hi) return 0; // non-monotonic times: no bracketing interval (the reference gives up after 16 rounds)
        g = min(max(g, lo), hi);
    }
    return 0;
}

// The interval entry held in registers; consecutive points almost always share it (points are time-sorted and an
// INS interval spans ~1 000-6 000 of them), which is checked per point, not assumed.
struct EntryRegs {
    double t_lo, t_hi, ts0, inv_dts;
    double A0, A1, A2, A3, A4, A5, A6, A7, A8;
    double kx, ky, kz, theta;
    double c0x, c0y, c0z, c1x, c1y, c1z;
};

__device__ __forceinline__ void load_entry(const Interval* __restrict__ it, EntryRegs& e) {
    const double2* __restrict__ p = reinterpret_cast<const double2*>(it);
    const double2 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3), v4 = __ldg(p + 4), v5 = __ldg(p + 5),
                  v6 = __ldg(p + 6), v7 = __ldg(p + 7), v8 = __ldg(p + 8), v9 = __ldg(p + 9), v10 = __ldg(p + 10),
                  v11 = __ldg(p + 11);
    e.t_lo = v0.x; e.t_hi = v0.y; e.ts0 = v1.x; e.inv_dts = v1.y;
    e.A0 = v2.x; e.A1 = v2.y; e.A2 = v3.x; e.A3 = v3.y; e.A4 = v4.x; e.A5 = v4.y; e.A6 = v5.x; e.A7 = v5.y; e.A8 = v6.x;
    e.kx = v6.y; e.ky = v7.x; e.kz = v7.y; e.theta = v8.x;
    e.c0x = v8.y; e.c0y = v9.x; e.c0z = v9.y; e.c1x = v10.x; e.c1y = v10.y; e.c1z = v11.x;
}

// does the cached interval entry (cur > 0) bracket time t, i.e. t_lo <= t < t_hi?  GPS seconds of week are positive,
// so IEEE doubles order like their bit patterns and the test is one unsigned 64-bit compare on the ALU pipe:
// (bits(t) - bits(t_lo)) < (bits(t_hi) - bits(t_lo)); a negative or NaN time wraps to a huge value and misses.
// With t_lo > t_hi > 0 (the fallback entry: t_lo = last sample, t_hi = first sample) the width wraps and the same
// compare is true exactly when NOT (t_hi <= t < t_lo), i.e. for the times the reference sends to the fallback pose
// (misc.cc:28-30), negative times included.
__device__ __forceinline__ bool key_brackets(unsigned long long lo, unsigned long long width, double t) {
    return ((unsigned long long) __double_as_longlong(t) - lo) < width;
}
__device__ __forceinline__ bool entry_brackets(const EntryRegs& e, double t) {
    const unsigned long long lo = (unsigned long long) __double_as_longlong(e.t_lo);
    const unsigned long long width = (unsigned long long) __double_as_longlong(e.t_hi) - lo;
    return key_brackets(lo, width, t);
}

// the same entry through plain loads: shared memory (k_deskew_bin stages the entries of a chunk next to its points) or,
// with a generic pointer, either space
__device__ __forceinline__ void load_entry_shared(const Interval* it, EntryRegs& e) {
    const double2* p = reinterpret_cast<const double2*>(it);
    const double2 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3], v4 = p[4], v5 = p[5], v6 = p[6], v7 = p[7], v8 = p[8], v9 = p[9],
                  v10 = p[10], v11 = p[11];
    e.t_lo = v0.x; e.t_hi = v0.y; e.ts0 = v1.x; e.inv_dts = v1.y;
    e.A0 = v2.x; e.A1 = v2.y; e.A2 = v3.x; e.A3 = v3.y; e.A4 = v4.x; e.A5 = v4.y; e.A6 = v5.x; e.A7 = v5.y; e.A8 = v6.x;
    e.kx = v6.y; e.ky = v7.x; e.kz = v7.y; e.theta = v8.x;
    e.c0x = v8.y; e.c0y = v9.x; e.c0z = v9.y; e.c1x = v10.x; e.c1y = v10.y; e.c1z = v11.x;
}

// Make `e` the entry for time t (misc.cc:25-60 semantics).  Fast when e already brackets t.  Otherwise the entry at the
// index predicted from the (near-uniform) INS spacing is loaded speculatively in one go and verified afterwards, so the
// common case costs one memory round trip, not two (probe, then load).
__device__ __forceinline__ void seek_entry(const KfView& kv, double t, EntryRegs& e, int& cur) {
    if (cur > 0 && e.t_lo > 0 && entry_brackets(e, t)) return;
    int idx = 0;
    if (!(kv.n_ins < 2 || kv.t0 > t || kv.t_last <= t)) { // misc.cc:28-30
        const int lo = 1, hi = kv.n_ins - 1;
        int g = 1 + (int) ((t - kv.t0) * kv.inv_mean_dt);
        g = min(max(g, lo), hi);
        if (g != cur) {
            load_entry(kv.tab + g, e);
            cur = g;
        }
        if ((e.t_lo <= t) && (t < e.t_hi)) return;
        idx = find_interval_slow(kv.tab, kv.n_ins, t, g, lo, hi);
    }
    if (idx != cur) {
        load_entry(kv.tab + idx, e);
        cur = idx;
    }
}

// The entry for time t straight from the keyframe's table in global memory (misc.cc:25-60 semantics), no cached state:
// the rare path of k_deskew_bin (points whose interval is not one of the two staged with their chunk).
__device__ __noinline__ const Interval* find_entry_global(const KfState* __restrict__ kf, const Interval* __restrict__ table, double t) {
    const Interval* __restrict__ tab = table + kf->table_begin;
    const int n = (int) kf->ins_count;
    int idx = 0;
    if (!(n < 2 || kf->ins_t0 > t || kf->ins_t_last <= t)) { // misc.cc:28-30
        int g = 1 + (int) ((t - kf->ins_t0) * kf->inv_mean_dt);
        g = min(max(g, 1), n - 1);
        const double2 ab = __ldg(reinterpret_cast<const double2*>(tab + g));
        idx = ((ab.x <= t) && (t < ab.y)) ? g : find_interval_slow(tab, n, t, g, 1, n - 1);
    }
    return tab + idx;
}

// sin(phi) and 1 - cos(phi) to full double precision for the small angles of one INS interval
__device__ __forceinline__ void sin_omc(double phi, double& sn, double& omc) {
    const double p2 = phi * phi;
    if (p2 < 2.44140625e-4) { // |phi| < 2^-6: truncation error < 1e-21 relative
        double a = fma(p2, -1.0 / 5040.0, 1.0 / 120.0);
        a = fma(p2, a, -1.0 / 6.0);
        sn = fma(phi * p2, a, phi);
        double b = fma(p2, -1.0 / 40320.0, 1.0 / 720.0);
        b = fma(p2, b, -1.0 / 24.0);
        b = fma(p2, b, 0.5);
        omc = p2 * b;
    } else if (p2 < 0.0625) { // |phi| < 0.25 rad: truncation error < 3e-18 relative
        double a = fma(p2, 1.0 / 6227020800.0, -1.0 / 39916800.0);
        a = fma(p2, a, 1.0 / 362880.0);
        a = fma(p2, a, -1.0 / 5040.0);
        a = fma(p2, a, 1.0 / 120.0);
        a = fma(p2, a, -1.0 / 6.0);
        sn = fma(phi * p2, a, phi);
        double b = fma(p2, 1.0 / 87178291200.0, -1.0 / 479001600.0);
        b = fma(p2, b, 1.0 / 3628800.0);
        b = fma(p2, b, -1.0 / 40320.0);
        b = fma(p2, b, 1.0 / 720.0);
        b = fma(p2, b, -1.0 / 24.0);
        b = fma(p2, b, 0.5);
        omc = p2 * b;
    } else {
        double s, cs, sh, ch;
        sincos(phi, &s, &cs);
        sincos(0.5 * phi, &sh, &ch);
        sn = s;
        omc = 2.0 * sh * sh;
    }
}

// tier-1 polynomials only (|phi| < 2^-6), straight-line: the caller checks theta of the interval once
__device__ __forceinline__ void sin_omc_small(double phi, double& sn, double& omc) {
    const double p2 = phi * phi;
    double a = fma(p2, -1.0 / 5040.0, 1.0 / 120.0);
    a = fma(p2, a, -1.0 / 6.0);
    sn = fma(phi * p2, a, phi);
    double b = fma(p2, -1.0 / 40320.0, 1.0 / 720.0);
    b = fma(p2, b, -1.0 / 24.0);
    b = fma(p2, b, 0.5);
    omc = p2 * b;
}

// Deskew one raw point into the keyframe's camera frame (le_vins.cc:455-461) with the entry in registers.
// The two f32 roundings of pointcloud.cc:68 are kept.
#ifndef LV_KO
#define LV_KO 0
#endif
#if LV_KO & 16 // timing experiment: no XU conversions (results are wrong)
#define LV_F2D(v) __hiloint2double(__float_as_int(v), 0)
#define LV_D2F(v) __int_as_float(__double2hiint(v))
#else
#define LV_F2D(v) ((double) (v))
#define LV_D2F(v) ((float) (v))
#endif
template <bool SMALL_ANGLE = false>
__device__ __forceinline__ void deskew_core(float x, float y, float z, double t, const EntryRegs& e, const Extrinsics& ex,
                                            float& ox, float& oy, float& oz) {
    // p_c1 = f32(R_cl p + t_cl)   (pointcloud.cc:38 -> 68, first rounding)
    const double xd = LV_F2D(x), yd = LV_F2D(y), zd = LV_F2D(z);
    const float p1x = LV_D2F(fma(ex.Rcl[6], zd, fma(ex.Rcl[3], yd, fma(ex.Rcl[0], xd, ex.tcl[0]))));
    const float p1y = LV_D2F(fma(ex.Rcl[7], zd, fma(ex.Rcl[4], yd, fma(ex.Rcl[1], xd, ex.tcl[1]))));
    const float p1z = LV_D2F(fma(ex.Rcl[8], zd, fma(ex.Rcl[5], yd, fma(ex.Rcl[2], xd, ex.tcl[2]))));
    const double ax = LV_F2D(p1x), ay = LV_F2D(p1y), az = LV_F2D(p1z);
    // u = R_bc p_c1 + t_bc
    const double ux = fma(ex.Rbc[6], az, fma(ex.Rbc[3], ay, fma(ex.Rbc[0], ax, ex.tbc[0])));
    const double uy = fma(ex.Rbc[7], az, fma(ex.Rbc[4], ay, fma(ex.Rbc[1], ax, ex.tbc[1])));
    const double uz = fma(ex.Rbc[8], az, fma(ex.Rbc[5], ay, fma(ex.Rbc[2], ax, ex.tbc[2])));
    const double s = (t - e.ts0) * e.inv_dts;
#if LV_KO & 64 // timing experiment: no rotation correction (28 of the 61 FP64 operations)
    const double wx = ux, wy = uy, wz = uz;
#else
    double sn, omc;
    if (SMALL_ANGLE)
        sin_omc_small(-s * e.theta, sn, omc);
    else
        sin_omc(-s * e.theta, sn, omc);
    // w = u + sin * (k x u) + (1 - cos) * k x (k x u)
    const double cx = fma(e.ky, uz, -(e.kz * uy)), cy = fma(e.kz, ux, -(e.kx * uz)), cz = fma(e.kx, uy, -(e.ky * ux));
    const double dx = fma(e.ky, cz, -(e.kz * cy)), dy = fma(e.kz, cx, -(e.kx * cz)), dz = fma(e.kx, cy, -(e.ky * cx));
    const double wx = fma(omc, dx, fma(sn, cx, ux));
    const double wy = fma(omc, dy, fma(sn, cy, uy));
    const double wz = fma(omc, dz, fma(sn, cz, uz));
#endif
    // p_c0 = A w + c0 + s c1   (second rounding, pointcloud.cc:45 -> 68)
    ox = LV_D2F(fma(e.A2, wz, fma(e.A1, wy, fma(e.A0, wx, fma(s, e.c1x, e.c0x)))));
    oy = LV_D2F(fma(e.A5, wz, fma(e.A4, wy, fma(e.A3, wx, fma(s, e.c1y, e.c0y)))));
    oz = LV_D2F(fma(e.A8, wz, fma(e.A7, wy, fma(e.A6, wx, fma(s, e.c1z, e.c0z)))));
}

// Range gate (le_vins.cc:447-451) on the squared norm (see Consts) + entry lookup + deskew.  false for gated points.
__device__ __forceinline__ bool deskew_point(float x, float y, float z, double time, const KfView& kv, const Extrinsics& ex,
                                             const Consts& c, EntryRegs& e, int& cur, float& ox, float& oy, float& oz) {
    const float n2 = v3f_sqnorm(x, y, z);
    if ((n2 > c.n2_far) || (n2 < c.n2_near)) return false;
    const double t = time + ex.td_b_l;
    seek_entry(kv, t, e, cur);
    deskew_core<false>(x, y, z, t, e, ex, ox, oy, oz);
    return true;
}

// ------------------------------------------------------------------------------------------------
// feature side: Camera::undistortPoints -> pixel2unitcam before the association, cam2world after it
// ------------------------------------------------------------------------------------------------
struct CameraModel {
    double fx, fy, cx, cy, skew;
    double k1, k2, p1, p2, k3; // OpenCV order of the first five distortion coefficients
};
// cv::undistortPoints(pts, pts, K, D, Mat(), K) (camera.cc:72-74): OpenCV's fixed-point iteration with its default
// criteria (5 iterations), R = identity, P = K, operation for operation (no contraction, IEEE division); f32 in, f64
// arithmetic, f32 out like cv::Point2f.
__device__ __forceinline__ void undistort_pixel(const CameraModel& c, float& pu, float& pv) {
    const double ifx = 1. / c.fx, ify = 1. / c.fy;
    const double u = (double) pu, v = (double) pv;
    double x = (u - c.cx) * ifx, y = (v - c.cy) * ify;
    const double x0 = x, y0 = y;
#pragma unroll 1
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = 1. / (1 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2); // numerator: k4 = k5 = k6 = 0
        if (icdist < 0) {
            x = (u - c.cx) * ifx;
            y = (v - c.cy) * ify;
            break;
        }
        const double deltaX = 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x);
        const double deltaY = c.p1 * (r2 + 2 * y * y) + 2 * c.p2 * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    pu = (float) (c.fx * x + c.skew * y + c.cx);
    pv = (float) (c.fy * y + c.cy);
}
// Camera::pixel2unitcam (camera.cc:123-135) then .cast<float>() (tracking.cc:391-393)
__device__ __forceinline__ void pixel_to_unit_ray(const CameraModel& c, float pu, float pv, float& qx, float& qy, float& qz) {
    const double y = ((double) pv - c.cy) / c.fy;
    const double x = ((double) pu - c.cx - c.skew * y) / c.fx;
    double r[3] = {x, y, 1.0};
    const double z = v3_dot(r, r);
    if (z > 0) {
        const double s = sqrt(z);
        r[0] /= s;
        r[1] /= s;
        r[2] /= s;
    }
    qx = (float) r[0];
    qy = (float) r[1];
    qz = (float) r[2];
}

// ------------------------------------------------------------------------------------------------
// Eigen::ColPivHouseholderQR<Matrix<double,5,3>> compute + solve (Eigen >= 3.3 algorithm, LAPACK-style norm down-dating)
// ------------------------------------------------------------------------------------------------
__device__ inline void colpiv_qr_solve_5x3(const double (*P)[3], double* x) {
    constexpr int ROWS = 5, COLS = 3;
    double q[COLS][ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++)
#pragma unroll
        for (int cI = 0; cI < COLS; cI++) q[cI][r] = P[r][cI];
    double hcoef[COLS];
    int perm[COLS] = {0, 1, 2};
    double nu[COLS], nd[COLS];
#pragma unroll
    for (int k = 0; k < COLS; k++) {
        double s = 0;
#pragma unroll
        for (int r = 0; r < ROWS; r++) s += q[k][r] * q[k][r];
        nd[k] = sqrt(s);
        nu[k] = nd[k];
    }
    double maxn = nu[0];
    if (nu[1] > maxn) maxn = nu[1];
    if (nu[2] > maxn) maxn = nu[2];
    const double th = maxn * DBL_EPSILON / (double) ROWS;
    const double threshold_helper = th * th;
    const double downdate = sqrt(DBL_EPSILON);
    int nonzero = COLS;

#pragma unroll
    for (int k = 0; k < COLS; k++) {
        int big = k;
        double bigv = nu[k];
#pragma unroll
        for (int j = k + 1; j < COLS; j++)
            if (nu[j] > bigv) {
                bigv = nu[j];
                big = j;
            }
        if (nonzero == COLS && bigv * bigv < threshold_helper * (double) (ROWS - k)) nonzero = k;
        if (k != big) {
#pragma unroll
            for (int j = k + 1; j < COLS; j++) {
                if (j == big) {
#pragma unroll
                    for (int r = 0; r < ROWS; r++) {
                        const double tmp = q[k][r];
                        q[k][r] = q[j][r];
                        q[j][r] = tmp;
                    }
                    double tmp = nu[k];
                    nu[k] = nu[j];
                    nu[j] = tmp;
                    tmp = nd[k];
                    nd[k] = nd[j];
                    nd[j] = tmp;
                    const int ti = perm[k];
                    perm[k] = perm[j];
                    perm[j] = ti;
                }
            }
        }
        const int len = ROWS - k;
        double tail_sq = 0;
#pragma unroll
        for (int r = 1; r < ROWS; r++)
            if (r < len) tail_sq += q[k][k + r] * q[k][k + r];
        const double c0 = q[k][k];
        double beta, tau;
        if (tail_sq <= DBL_MIN) {
            tau = 0;
            beta = c0;
#pragma unroll
            for (int r = 1; r < ROWS; r++)
                if (r < len) q[k][k + r] = 0;
        } else {
            beta = sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
#pragma unroll
            for (int r = 1; r < ROWS; r++)
                if (r < len) q[k][k + r] = q[k][k + r] / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hcoef[k] = tau;
        q[k][k] = beta;
        if (tau != 0) {
#pragma unroll
            for (int j = k + 1; j < COLS; j++) {
                double tmp = 0;
#pragma unroll
                for (int r = 1; r < ROWS; r++)
                    if (r < len) tmp += q[k][k + r] * q[j][k + r];
                tmp += q[j][k];
                q[j][k] -= tau * tmp;
#pragma unroll
                for (int r = 1; r < ROWS; r++)
                    if (r < len) q[j][k + r] -= (tau * q[k][k + r]) * tmp;
            }
        }
#pragma unroll
        for (int j = k + 1; j < COLS; j++) {
            if (nu[j] != 0) {
                double temp = fabs(q[j][k]) / nu[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0 ? 0 : temp;
                const double ratio = nu[j] / nd[j];
                const double temp2 = temp * (ratio * ratio);
                if (temp2 <= downdate) {
                    double s = 0;
#pragma unroll
                    for (int r = 1; r < ROWS; r++)
                        if (r >= k + 1) s += q[j][r] * q[j][r];
                    nd[j] = sqrt(s);
                    nu[j] = nd[j];
                } else {
                    nu[j] *= sqrt(temp);
                }
            }
        }
    }

    double c[ROWS] = {-1.0, -1.0, -1.0, -1.0, -1.0}; // mb = -Ones (tracking.cc:585)
#pragma unroll
    for (int k = 0; k < COLS; k++) {
        if (k < nonzero) {
            const double tau = hcoef[k];
            const int len = ROWS - k;
            if (tau != 0) {
                double tmp = 0;
#pragma unroll
                for (int r = 1; r < ROWS; r++)
                    if (r < len) tmp += q[k][k + r] * c[k + r];
                tmp += c[k];
                c[k] -= tau * tmp;
#pragma unroll
                for (int r = 1; r < ROWS; r++)
                    if (r < len) c[k + r] -= (tau * q[k][k + r]) * tmp;
            }
        }
    }
#pragma unroll
    for (int i = COLS - 1; i >= 0; i--) {
        if (i < nonzero) {
            c[i] /= q[i][i];
#pragma unroll
            for (int r = 0; r < COLS; r++)
                if (r < i) c[r] -= c[i] * q[i][r];
        }
    }
    x[0] = x[1] = x[2] = 0;
#pragma unroll
    for (int i = 0; i < COLS; i++)
        if (i < nonzero) {
            // x[perm[i]] = c[i] without dynamic register indexing
            if (perm[i] == 0) x[0] = c[i];
            if (perm[i] == 1) x[1] = c[i];
            if (perm[i] == 2) x[2] = c[i];
        }
}

// ------------------------------------------------------------------------------------------------
// symmetric 3x3 eigen-decomposition (cyclic Jacobi): stands in for the Eigen::EigenSolver call on the neighbourhood
// covariance at pointcloud.cc:132-134.  w: eigenvalues, v: eigenvectors as columns (row-major).
// ------------------------------------------------------------------------------------------------
__device__ inline void sym_eigen3(const double* a_in, double* w, double* v) {
    double a[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a[i][j] = a_in[3 * i + j];
    for (int sweep = 0; sweep < 60; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; k++) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = cs * akp - sn * akq;
                    a[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; k++) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = cs * apk - sn * aqk;
                    a[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = cs * vkp - sn * vkq;
                    V[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    for (int i = 0; i < 3; i++) {
        w[i] = a[i][i];
        for (int j = 0; j < 3; j++) v[3 * i + j] = V[i][j];
    }
}

} // namespace lv
