/*
 * levins_oracle.c -- CPU ORACLE (test infrastructure, see levins_oracle.h).
 *
 * Restates, function by function, the reference's CPU algorithm for the LiDAR
 * depth-enhancement path.  "parity unpinned": the reference has no golden vectors and
 * cannot be built in this image; see the header for how the oracle is pinned instead.
 *
 * Build: gcc -O3 -ffp-contract=off -fopenmp (no -march, like the reference's
 * CMakeLists.txt:7-10 => SSE2 baseline, no FMA contraction, FLT_EVAL_METHOD 0).
 */
#include "levins_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * constants -- tracking.h:119-135 evaluated with the same double expressions
 * ------------------------------------------------------------------------------------------ */
static double k_angle(void) { return 0.3 * M_PI / 180.0; }     /* ASSOCIATE_RANGE_IMAGE_ANGLE  tracking.h:119 */
static double k_fov(void) { return 90.0 * M_PI / 180.0; }      /* ASSOCIATE_IMAGE_FOV          tracking.h:124 */
double orc_const_range_image_angle(void) { return k_angle(); }
double orc_const_tan_value(void) { return tan(k_fov() * 0.5); } /* tracking.h:128 */
double orc_const_nearest_sq_threshold(void) {                   /* tracking.h:130-133 */
    double thr = sin(k_angle()) * 3;
    return pow(thr, 2);
}
int orc_const_image_size(void) { return (int) (k_fov() / k_angle()); } /* tracking.h:126 */

static const double NEAREST_DEPTH = 1;    /* visual_mappoint.h:50 */
static const double FARTHEST_DEPTH = 200; /* visual_mappoint.h:51 */

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * small fixed-size algebra with Eigen's reduction orders (SURVEY 8a note 4):
 *   Vector3f redux: e0 + (e1 + e2)        Vector3d redux (SSE2 packets): (e0 + e1) + e2
 * ------------------------------------------------------------------------------------------ */
static inline float v3f_sqnorm(float x, float y, float z) { return x * x + (y * y + z * z); }
static inline float v3f_dot(const float a[3], const float b[3]) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
static inline double v3d_dot(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline double v3d_norm(const double a[3]) { return sqrt(v3d_dot(a, a)); }

/* y = R x (R column-major), sum order ((c0 x0) + c1 x1) + c2 x2 */
static inline void m3_mul_v(const double R[9], const double x[3], double y[3]) {
    for (int i = 0; i < 3; i++) y[i] = (R[i] * x[0] + R[3 + i] * x[1]) + R[6 + i] * x[2];
}
/* y = R^T x */
static inline void m3t_mul_v(const double R[9], const double x[3], double y[3]) {
    for (int i = 0; i < 3; i++) y[i] = (R[3 * i] * x[0] + R[3 * i + 1] * x[1]) + R[3 * i + 2] * x[2];
}
/* C = A B */
static inline void m3_mul_m(const double A[9], const double B[9], double C[9]) {
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) C[3 * j + i] = (A[i] * B[3 * j] + A[3 + i] * B[3 * j + 1]) + A[6 + i] * B[3 * j + 2];
}
/* C = A^T B */
static inline void m3t_mul_m(const double A[9], const double B[9], double C[9]) {
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++)
            C[3 * j + i] = (A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1]) + A[3 * i + 2] * B[3 * j + 2];
}

/* Eigen generic quat_product (Quaternion.h), q = (x, y, z, w) */
static inline void quat_mul(const double a[4], const double b[4], double r[4]) {
    double ax = a[0], ay = a[1], az = a[2], aw = a[3];
    double bx = b[0], by = b[1], bz = b[2], bw = b[3];
    r[3] = aw * bw - ax * bx - ay * by - az * bz;
    r[0] = aw * bx + ax * bw + ay * bz - az * by;
    r[1] = aw * by + ay * bw + az * bx - ax * bz;
    r[2] = aw * bz + az * bw + ax * by - ay * bx;
}
static inline double quat_sqnorm(const double q[4]) { return ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]; }
/* Eigen QuaternionBase::inverse(): conjugate / squaredNorm, zero quaternion if n2 == 0 */
static inline void quat_inverse(const double q[4], double r[4]) {
    double n2 = quat_sqnorm(q);
    if (n2 > 0) {
        r[0] = -q[0] / n2;
        r[1] = -q[1] / n2;
        r[2] = -q[2] / n2;
        r[3] = q[3] / n2;
    } else {
        r[0] = r[1] = r[2] = r[3] = 0;
    }
}
/* Eigen normalize(): z = squaredNorm; if (z > 0) coeffs /= sqrt(z) */
static inline void quat_normalize(double q[4]) {
    double z = quat_sqnorm(q);
    if (z > 0) {
        double s = sqrt(z);
        q[0] /= s;
        q[1] /= s;
        q[2] /= s;
        q[3] /= s;
    }
}

/* Eigen QuaternionBase::toRotationMatrix() */
void orc_quat_to_matrix(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); /* (0,0) */
    R[3] = txy - twz;       /* (0,1) */
    R[6] = txz + twy;       /* (0,2) */
    R[1] = txy + twz;       /* (1,0) */
    R[4] = 1 - (txx + tzz); /* (1,1) */
    R[7] = tyz - twx;       /* (1,2) */
    R[2] = txz - twy;       /* (2,0) */
    R[5] = tyz + twx;       /* (2,1) */
    R[8] = 1 - (txx + tyy); /* (2,2) */
}

/* rotation.cc:65-68: Eigen::AngleAxisd(quaternion); return angle * axis.
 * AngleAxis::operator=(QuaternionBase): n = vec.norm(); if n < eps: n = stableNorm();
 * if n != 0: angle = 2 atan2(n, |w|); if w < 0: n = -n; axis = vec / n; else angle 0, axis (1,0,0). */
void orc_quaternion2vector(const double q[4], double rvec[3]) {
    double n = v3d_norm(q);
    if (n < DBL_EPSILON) {
        /* stableNorm(): scale by the max |coeff| before summing (same value to the last ulp or two) */
        double m = fmax(fabs(q[0]), fmax(fabs(q[1]), fabs(q[2])));
        if (m > 0) {
            double a = q[0] / m, b = q[1] / m, c = q[2] / m;
            n = m * sqrt((a * a + b * b) + c * c);
        } else {
            n = 0;
        }
    }
    double angle, axis[3];
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
    rvec[0] = angle * axis[0];
    rvec[1] = angle * axis[1];
    rvec[2] = angle * axis[2];
}

/* rotation.cc:59-63: angle = rotvec.norm(); vec = rotvec.normalized(); Quaterniond(AngleAxisd(angle, vec)).
 * normalized(): z = squaredNorm; z > 0 ? v / sqrt(z) : v.   Quaternion(AngleAxis): ha = 0.5 angle; w = cos, vec = sin * axis. */
void orc_rotvec2quaternion(const double rvec[3], double q[4]) {
    double angle = v3d_norm(rvec);
    double z = v3d_dot(rvec, rvec);
    double vec[3] = {rvec[0], rvec[1], rvec[2]};
    if (z > 0) {
        double s = sqrt(z);
        vec[0] /= s;
        vec[1] /= s;
        vec[2] /= s;
    }
    double ha = 0.5 * angle;
    double sn = sin(ha);
    q[3] = cos(ha);
    q[0] = sn * vec[0];
    q[1] = sn * vec[1];
    q[2] = sn * vec[2];
}

/* ------------------------------------------------------------------------------------------
 * misc.cc:25-60  MISC::getInsWindowIndex
 * ------------------------------------------------------------------------------------------ */
size_t orc_get_ins_window_index(const orc_ins *w, size_t n, double time) {
    if (n == 0 || (w[0].time > time) || (w[n - 1].time <= time)) {
        return 0; /* misc.cc:28-30 */
    }
    size_t index = 0;
    size_t sta = 0, end = n;
    int counts = 0;
    while (1) { /* misc.cc:37-57 */
        size_t mid = (sta + end) / 2;
        double first = w[mid - 1].time;
        double second = w[mid].time;
        if ((first <= time) && (time < second)) {
            index = mid;
            break;
        } else if (first > time) {
            end = mid;
        } else if (second <= time) {
            sta = mid;
        }
        if (counts++ > 15) { /* misc.cc:53-56 */
            break;
        }
    }
    return index;
}

/* misc.cc:80-95  MISC::statePoseInterpolation */
void orc_state_pose_interpolation(const orc_ins *s0, const orc_ins *s1, double midtime, double p[3], double q[4]) {
    double dp[3] = {s1->p[0] - s0->p[0], s1->p[1] - s0->p[1], s1->p[2] - s0->p[2]}; /* :84 */
    double q1inv[4], dq[4], rvec[3];
    quat_inverse(s1->q, q1inv);
    quat_mul(q1inv, s0->q, dq);     /* :85 */
    orc_quaternion2vector(dq, rvec); /* :86 */

    double scale = (midtime - s0->state_time) / (s1->state_time - s0->state_time); /* :88 */
    rvec[0] = rvec[0] * scale;                                                     /* :89 */
    rvec[1] = rvec[1] * scale;
    rvec[2] = rvec[2] * scale;
    orc_rotvec2quaternion(rvec, dq); /* :90 */

    p[0] = s0->p[0] + dp[0] * scale; /* :92 */
    p[1] = s0->p[1] + dp[1] * scale;
    p[2] = s0->p[2] + dp[2] * scale;
    double dqinv[4];
    quat_inverse(dq, dqinv);
    quat_mul(s0->q, dqinv, q); /* :93 */
    quat_normalize(q);         /* :94 */
}

/* misc.cc:97-103  MISC::stateToPoseWithExtrinsic */
void orc_state_to_pose_with_extrinsic(const double p[3], const double q[4], const orc_pose *ext, orc_pose *pose) {
    double Rq[9], Rt[3];
    orc_quat_to_matrix(q, Rq);
    m3_mul_v(Rq, ext->t, Rt);
    pose->t[0] = p[0] + Rt[0]; /* :100 */
    pose->t[1] = p[1] + Rt[1];
    pose->t[2] = p[2] + Rt[2];
    m3_mul_m(Rq, ext->R, pose->R); /* :101 */
}

/* misc.cc:62-78  MISC::getPoseFromInsWindow */
int orc_get_pose_from_ins_window(const orc_ins *w, size_t n, const orc_pose *ext, double time, orc_pose *pose) {
    size_t index = orc_get_ins_window_index(w, n, time);
    if (index > 0) {
        double p[3], q[4];
        orc_state_pose_interpolation(&w[index - 1], &w[index], time, p, q);
        orc_state_to_pose_with_extrinsic(p, q, ext, pose);
        return 1;
    } else {
        orc_state_to_pose_with_extrinsic(w[n - 1].p, w[n - 1].q, ext, pose); /* :75 window.back() */
        return 0;
    }
}

/* ------------------------------------------------------------------------------------------
 * pointcloud.cc
 * ------------------------------------------------------------------------------------------ */
/* pointcloud.cc:65-71: projected = point; xyz = (R * xyz.cast<double>() + t).cast<float>() */
void orc_point_projection(const orc_pose *pose, const orc_point_xyzi *in, orc_point_xyzi *out) {
    double v[3] = {(double) in->x, (double) in->y, (double) in->z}, r[3];
    m3_mul_v(pose->R, v, r);
    orc_point_xyzi o = *in;
    o.x = (float) (r[0] + pose->t[0]);
    o.y = (float) (r[1] + pose->t[1]);
    o.z = (float) (r[2] + pose->t[2]);
    *out = o;
}

static void relative_pose(const orc_pose *pose0, const orc_pose *pose1, orc_pose *rel) {
    double dt[3] = {pose1->t[0] - pose0->t[0], pose1->t[1] - pose0->t[1], pose1->t[2] - pose0->t[2]};
    m3t_mul_m(pose0->R, pose1->R, rel->R); /* pointcloud.cc:30, 42 */
    m3t_mul_v(pose0->R, dt, rel->t);       /* pointcloud.cc:31, 43 */
}

/* pointcloud.cc:28-34 */
void orc_relative_projection(const orc_pose *pose0, const orc_pose *pose1, const orc_point_xyzi *in,
                             orc_point_xyzi *out) {
    orc_pose rel;
    relative_pose(pose0, pose1, &rel);
    orc_point_projection(&rel, in, out);
}

/* pointcloud.cc:36-48 */
void orc_relative_projection_to_camera(const orc_pose *pose_c_l, const orc_pose *pose_cam0, const orc_pose *pose_cam1,
                                       const orc_point_xyzi *in, orc_point_xyzi *out) {
    orc_point_xyzi cam1;
    orc_point_projection(pose_c_l, in, &cam1); /* :38, first f32 rounding */
    orc_pose rel;
    relative_pose(pose_cam0, pose_cam1, &rel); /* :41-43 */
    orc_point_projection(&rel, &cam1, out);    /* :45, second f32 rounding */
}

/* pointcloud.cc:50-63 (tbb::parallel_for -> OpenMP) */
void orc_point_cloud_projection(const orc_pose *pose, const orc_point_xyzi *src, orc_point_xyzi *dst, size_t n,
                                int nthreads) {
    (void) nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (long k = 0; k < (long) n; k++) orc_point_projection(pose, &src[k], &dst[k]);
}

/* pointcloud.cc:107-109: (a - b).squaredNorm() in f32, widened on return */
double orc_point_square_distance(const orc_point_xyzi *a, const orc_point_xyzi *b) {
    return (double) v3f_sqnorm(a->x - b->x, a->y - b->y, a->z - b->z);
}

/* ------------------------------------------------------------------------------------------
 * ring buffer, le_vins.cc:149-163 and 413-424  (std::deque stand-in: flat array + head)
 * ------------------------------------------------------------------------------------------ */
struct orc_ring {
    orc_point_xyzit *data;
    size_t head, size, cap;
};
orc_ring *orc_ring_create(void) { return (orc_ring *) calloc(1, sizeof(orc_ring)); }
void orc_ring_destroy(orc_ring *r) {
    if (r) {
        free(r->data);
        free(r);
    }
}
size_t orc_ring_size(const orc_ring *r) { return r->size; }
const orc_point_xyzit *orc_ring_data(const orc_ring *r) { return r->data + r->head; }
void orc_ring_append(orc_ring *r, const orc_point_xyzit *pts, size_t n) {
    if (n == 0 && r->size == 0) return; /* the reference would read back() of an empty deque: UB; define no-op */
    if (r->head + r->size + n > r->cap) {
        size_t ncap = (r->size + n) * 2 + 1024;
        orc_point_xyzit *nd = (orc_point_xyzit *) malloc(ncap * sizeof(orc_point_xyzit));
        if (r->size) memcpy(nd, r->data + r->head, r->size * sizeof(orc_point_xyzit));
        free(r->data);
        r->data = nd;
        r->head = 0;
        r->cap = ncap;
    }
    memcpy(r->data + r->head + r->size, pts, n * sizeof(orc_point_xyzit)); /* le_vins.cc:155 */
    r->size += n;
    const double reserved_time = 10.0 + 0.5 * 2; /* le_vins.cc:156; RESERVED_BUFFER_LENGTH=10, STACK_INTERVAL=0.5 */
    while (fabs(r->data[r->head + r->size - 1].time - r->data[r->head].time) > reserved_time) { /* :157-159 */
        r->head++;
        r->size--;
    }
}
size_t orc_ring_evict_before(orc_ring *r, double last_stamp) {
    size_t popped = 0;
    while (r->size && r->data[r->head].time < last_stamp) { /* le_vins.cc:413-424 */
        r->head++;
        r->size--;
        popped++;
    }
    return popped;
}

/* ------------------------------------------------------------------------------------------
 * le_vins.cc:392-470  VINS::pointcloudStackForVisualTracking  (window logic + deskew lambda)
 * ------------------------------------------------------------------------------------------ */
int orc_stack_deskew(const orc_point_xyzit *buf, size_t n_buf, double stamp, double stack_interval,
                     const orc_ins *ins, size_t n_ins, const orc_pose *pose_b_c, const orc_pose *pose_c_l,
                     double td_b_l, size_t *first_out, size_t *n_out, orc_point_xyzi *out, int32_t *ins_index,
                     uint8_t *kept, int nthreads) {
    double last_stamp = stamp - stack_interval; /* :395 */
    double curr_stamp = stamp;

    size_t first = 0; /* :413-424 pop_front while time < last_stamp */
    while (first < n_buf && buf[first].time < last_stamp) first++;
    *first_out = first;
    if (first == n_buf) { /* :426-429 */
        *n_out = 0;
        return -1;
    }
    const orc_point_xyzit *pts = buf + first;
    size_t size = n_buf - first;

    size_t num_projected = size; /* :431-437 */
    for (size_t k = 0; k < size; k++) {
        if (pts[k].time > curr_stamp) {
            num_projected = k;
            break;
        }
    }
    *n_out = num_projected;

    orc_pose last_keyframe_pose; /* :441-442 */
    orc_get_pose_from_ins_window(ins, n_ins, pose_b_c, curr_stamp, &last_keyframe_pose);

    (void) nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (long k = 0; k < (long) num_projected; k++) { /* :445-464 */
        orc_point_xyzi zero = {0, 0, 0, 1.0f, 0, {0, 0, 0}}; /* resize() default point (:438) */
        const orc_point_xyzit *raw = &pts[k];
        float norm = sqrtf(v3f_sqnorm(raw->x, raw->y, raw->z)); /* :447 */
        if ((norm > FARTHEST_DEPTH) || (norm < NEAREST_DEPTH)) { /* :448-451 */
            out[k] = zero;
            if (ins_index) ins_index[k] = -1;
            if (kept) kept[k] = 0;
            continue;
        }
        double t = raw->time + td_b_l; /* :455 */
        orc_pose pose;
        orc_get_pose_from_ins_window(ins, n_ins, pose_b_c, t, &pose);
        if (ins_index) ins_index[k] = (int32_t) orc_get_ins_window_index(ins, n_ins, t);
        if (kept) kept[k] = 1;

        orc_point_xyzi point = zero; /* :458-459 */
        point.x = raw->x;
        point.y = raw->y;
        point.z = raw->z;
        orc_relative_projection_to_camera(pose_c_l, &last_keyframe_pose, &pose, &point, &out[k]); /* :460-461 */
    }
    return num_projected != 0; /* :469 */
}

/* ------------------------------------------------------------------------------------------
 * fdlibm-lineage atanf / atan2f (glibc <= 2.40 sysdeps/ieee754/flt-32/{s_atanf,e_atan2f}.c algorithm,
 * restated).  The GPU kernel uses the same operation sequence; tests check it equals this image's libm.
 * ------------------------------------------------------------------------------------------ */
static inline int32_t f2i(float x) {
    int32_t i;
    memcpy(&i, &x, 4);
    return i;
}
static inline float i2f(int32_t i) {
    float x;
    memcpy(&x, &i, 4);
    return x;
}
static const float AT_HI[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
static const float AT_LO[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
static const float AT_C[11] = {3.3333334327e-01f,  -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                               9.0908870101e-02f,  -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                               4.9768779427e-02f,  -3.6531571299e-02f, 1.6285819933e-02f};
static float atanf_fdlibm(float x) {
    float w, s1, s2, z;
    int32_t ix, hx, id;
    hx = f2i(x);
    ix = hx & 0x7fffffff;
    if (ix >= 0x4c000000) { /* |x| >= 2^25 */
        if (ix > 0x7f800000) return x + x;
        if (hx > 0) return AT_HI[3] + AT_LO[3];
        return -AT_HI[3] - AT_LO[3];
    }
    if (ix < 0x3ee00000) {              /* |x| < 0.4375 */
        if (ix < 0x31000000) return x; /* |x| < 2^-29 */
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {     /* |x| < 1.1875 */
            if (ix < 0x3f300000) { /* 7/16 <= |x| < 11/16 */
                id = 0;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else {
                id = 1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000) { /* |x| < 2.4375 */
                id = 2;
                x = (x - 1.5f) / (1.0f + 1.5f * x);
            } else {
                id = 3;
                x = -1.0f / x;
            }
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (AT_C[0] + w * (AT_C[2] + w * (AT_C[4] + w * (AT_C[6] + w * (AT_C[8] + w * AT_C[10])))));
    s2 = w * (AT_C[1] + w * (AT_C[3] + w * (AT_C[5] + w * (AT_C[7] + w * AT_C[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = AT_HI[id] - ((x * (s1 + s2) - AT_LO[id]) - x);
    return (hx < 0) ? -z : z;
}
float orc_atan2f_fdlibm(float y, float x) {
    static const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
                       pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    float z;
    int32_t k, m, hx, hy, ix, iy;
    hx = f2i(x);
    ix = hx & 0x7fffffff;
    hy = f2i(y);
    iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanf_fdlibm(y);
    m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
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
    k = (iy - ix) >> 23;
    if (k > 60)
        z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60)
        z = 0.0f;
    else
        z = atanf_fdlibm(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return i2f(f2i(z) ^ (int32_t) 0x80000000);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

/* compares the restatement with this machine's libm atan2f on n pseudo-random inputs (finite coordinates and raw
 * bit patterns); returns the number of mismatching results (NaN == NaN). */
long orc_atan2f_selfcheck(long n, unsigned long long seed) {
    unsigned long long s = seed ? seed : 88172645463325252ULL;
    long bad = 0;
    for (long i = 0; i < n; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        float y = (float) ((long long) (s & 0xffffff) - 0x800000) * (200.0f / 0x800000);
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        float x = (float) ((long long) (s & 0xffffff) - 0x800000) * (200.0f / 0x800000);
        if (i % 7 == 0) {
            x = i2f((int32_t) (s >> 20));
            y = i2f((int32_t) (s >> 3));
        }
        float a = atan2f(y, x), b = orc_atan2f_fdlibm(y, x);
        if (f2i(a) != f2i(b) && !(a != a && b != b)) bad++;
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------
 * tracking.cc:298-358  Tracking::depthAssociationProjection  (binning + compaction)
 * ------------------------------------------------------------------------------------------ */
size_t orc_range_image(const orc_point_xyzi *cloud, size_t n, int libm_mode, int32_t *cell_of_point, int32_t *winner,
                       orc_point_xyzi *unit_sphere, int32_t *cell_to_compact) {
    const double ANGLE = k_angle();
    const double FOV = k_fov();
    const int SIZE = orc_const_image_size();
    const double TAN = orc_const_tan_value();

    float *range_image = (float *) malloc(sizeof(float) * SIZE * SIZE); /* :305 */
    for (int i = 0; i < SIZE * SIZE; i++) {
        range_image[i] = FLT_MAX;
        winner[i] = -1;
    }

    for (size_t k = 0; k < n; k++) { /* :307-336 */
        const orc_point_xyzi *point = &cloud[k];
        if (cell_of_point) cell_of_point[k] = -1;

        /* :311-314. x/z is a float division; fabs on a float returns the same value under both
         * overload resolutions; the compare promotes to double. */
        if ((point->z < 0) || ((double) fabsf(point->x / point->z) > TAN) ||
            ((double) fabsf(point->y / point->z) > TAN)) {
            continue;
        }

        double row_angle, col_angle;
        float s = point->x * point->x + point->z * point->z; /* float products and sum under both resolutions */
        if (libm_mode == ORC_LIBM_FLOAT) {
            /* std::sqrt(float), std::atan2(float, float): float results, then widened by the + double */
            row_angle = (double) atan2f(point->y, sqrtf(s)) + (M_PI * 0.5 - FOV * 0.5); /* :317-318 */
            col_angle = (double) atan2f(point->z, point->x) - (M_PI * 0.5 - FOV * 0.5); /* :320 */
        } else {
            row_angle = atan2((double) point->y, sqrt((double) s)) + (M_PI * 0.5 - FOV * 0.5);
            col_angle = atan2((double) point->z, (double) point->x) - (M_PI * 0.5 - FOV * 0.5);
        }

        int row = (int) round(row_angle / ANGLE); /* :323 */
        int col = (int) round(col_angle / ANGLE); /* :324 */

        if ((row < 0) || (col < 0) || (row >= SIZE) || (col >= SIZE)) { /* :326-328 */
            continue;
        }

        float distance = sqrtf(v3f_sqnorm(point->x, point->y, point->z)); /* :330 */
        int index = col + row * SIZE;                                     /* :331 */
        if (cell_of_point) cell_of_point[k] = index;
        if (distance < range_image[index]) { /* :332-335 */
            range_image[index] = distance;
            winner[index] = (int32_t) k;
        }
    }

    size_t m = 0; /* :340-354 */
    for (int i = 0; i < SIZE; i++) {
        for (int j = 0; j < SIZE; j++) {
            int index = j + i * SIZE;
            if (cell_to_compact) cell_to_compact[index] = -1;
            if (range_image[index] != FLT_MAX) {
                const orc_point_xyzi *v = &cloud[winner[index]];
                orc_point_xyzi p = {0, 0, 0, 1.0f, 0, {0, 0, 0}};
                float n2 = v3f_sqnorm(v->x, v->y, v->z);
                p.intensity = sqrtf(n2); /* :348 */
                if (n2 > 0) {            /* :349 normalized() */
                    float nn = sqrtf(n2);
                    p.x = v->x / nn;
                    p.y = v->y / nn;
                    p.z = v->z / nn;
                } else {
                    p.x = v->x;
                    p.y = v->y;
                    p.z = v->z;
                }
                if (cell_to_compact) cell_to_compact[index] = (int32_t) m;
                unit_sphere[m++] = p;
            }
        }
    }
    free(range_image);
    return m;
}

/* ------------------------------------------------------------------------------------------
 * exact k-NN.  FLANN L2_Simple<float>: result = 0; for each dim: diff = a - b; result += diff*diff
 * ------------------------------------------------------------------------------------------ */
static inline float l2_simple(const float a[3], const float *b) {
    float r = 0;
    float d0 = a[0] - b[0];
    r += d0 * d0;
    float d1 = a[1] - b[1];
    r += d1 * d1;
    float d2 = a[2] - b[2];
    r += d2 * d2;
    return r;
}

/* sorted insertion ordered by (d2, index) -- the build's tie rule (SURVEY 8a note 5) */
static inline void knn_insert(int k, int *count, int32_t *idx, float *d2, float d, int32_t i) {
    if (*count == k) {
        if (d > d2[k - 1] || (d == d2[k - 1] && i > idx[k - 1])) return;
    }
    int pos = (*count < k) ? *count : k - 1;
    while (pos > 0 && (d2[pos - 1] > d || (d2[pos - 1] == d && idx[pos - 1] > i))) {
        d2[pos] = d2[pos - 1];
        idx[pos] = idx[pos - 1];
        pos--;
    }
    d2[pos] = d;
    idx[pos] = i;
    if (*count < k) (*count)++;
}

int orc_brute_knn(const orc_point_xyzi *pts, size_t m, const float q[3], int k, int32_t *idx, float *d2) {
    int count = 0;
    for (size_t i = 0; i < m; i++) {
        float d = l2_simple(q, &pts[i].x);
        knn_insert(k, &count, idx, d2, d, (int32_t) i);
    }
    return count;
}

/* FLANN KDTreeSingleIndex restatement (kdtree_single_index.h): bounding-box tree, middleSplit_,
 * leaf_max_size 15, reorder = true, exact search (eps = 0).  Used for the CPU-baseline timing and
 * cross-checked against the brute-force search; both are exact so they agree up to distance ties. */
typedef struct kd_node {
    int left, right;      /* leaf: point range */
    int divfeat;          /* -1 for leaf */
    float divlow, divhigh;
    int child1, child2;
} kd_node;
struct orc_kdtree {
    size_t m;
    float *data;   /* reordered, 3 floats per point */
    int32_t *vind; /* original index per reordered slot */
    kd_node *nodes;
    int n_nodes, cap_nodes;
    float bb_lo[3], bb_hi[3];
};
#define KD_LEAF 15

static int kd_new_node(orc_kdtree *t) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
        t->nodes = (kd_node *) realloc(t->nodes, sizeof(kd_node) * t->cap_nodes);
    }
    return t->n_nodes++;
}

static void kd_middle_split(const float *pts, int32_t *ind, int count, int *index, int *cutfeat, float *cutval,
                            const float lo[3], const float hi[3]) {
    const float EPS = 0.00001f;
    float max_span = hi[0] - lo[0];
    for (int i = 1; i < 3; i++) {
        float span = hi[i] - lo[i];
        if (span > max_span) max_span = span;
    }
    float max_spread = -1;
    *cutfeat = 0;
    for (int i = 0; i < 3; i++) {
        float span = hi[i] - lo[i];
        if (span > (1 - EPS) * max_span) {
            float mn = pts[3 * ind[0] + i], mx = mn;
            for (int k = 1; k < count; k++) {
                float v = pts[3 * ind[k] + i];
                if (v < mn) mn = v;
                if (v > mx) mx = v;
            }
            float spread = mx - mn;
            if (spread > max_spread) {
                *cutfeat = i;
                max_spread = spread;
            }
        }
    }
    int cf = *cutfeat;
    float split_val = (lo[cf] + hi[cf]) / 2;
    float mn = pts[3 * ind[0] + cf], mx = mn;
    for (int k = 1; k < count; k++) {
        float v = pts[3 * ind[k] + cf];
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
    if (split_val < mn)
        *cutval = mn;
    else if (split_val > mx)
        *cutval = mx;
    else
        *cutval = split_val;

    /* planeSplit: ind[0..lim1) < cutval, ind[lim1..lim2) == cutval, ind[lim2..count) > cutval */
    int left = 0, right = count - 1;
    for (;;) {
        while (left <= right && pts[3 * ind[left] + cf] < *cutval) ++left;
        while (left <= right && pts[3 * ind[right] + cf] >= *cutval) --right;
        if (left > right) break;
        int32_t tmp = ind[left];
        ind[left] = ind[right];
        ind[right] = tmp;
        ++left;
        --right;
    }
    int lim1 = left;
    right = count - 1;
    for (;;) {
        while (left <= right && pts[3 * ind[left] + cf] <= *cutval) ++left;
        while (left <= right && pts[3 * ind[right] + cf] > *cutval) --right;
        if (left > right) break;
        int32_t tmp = ind[left];
        ind[left] = ind[right];
        ind[right] = tmp;
        ++left;
        --right;
    }
    int lim2 = left;
    if (lim1 > count / 2)
        *index = lim1;
    else if (lim2 < count / 2)
        *index = lim2;
    else
        *index = count / 2;
}

static int kd_divide(orc_kdtree *t, const float *pts, int32_t *vind, int left, int right, float lo[3], float hi[3]) {
    int ni = kd_new_node(t);
    if ((right - left) <= KD_LEAF) {
        kd_node nd;
        nd.child1 = nd.child2 = -1;
        nd.divfeat = -1;
        nd.left = left;
        nd.right = right;
        nd.divlow = nd.divhigh = 0;
        for (int i = 0; i < 3; i++) {
            lo[i] = hi[i] = pts[3 * vind[left] + i];
        }
        for (int k = left + 1; k < right; k++)
            for (int i = 0; i < 3; i++) {
                float v = pts[3 * vind[k] + i];
                if (v < lo[i]) lo[i] = v;
                if (v > hi[i]) hi[i] = v;
            }
        t->nodes[ni] = nd;
        return ni;
    }
    int idx, cutfeat;
    float cutval;
    kd_middle_split(pts, vind + left, right - left, &idx, &cutfeat, &cutval, lo, hi);
    float llo[3], lhi[3], rlo[3], rhi[3];
    memcpy(llo, lo, sizeof(llo));
    memcpy(lhi, hi, sizeof(lhi));
    memcpy(rlo, lo, sizeof(rlo));
    memcpy(rhi, hi, sizeof(rhi));
    lhi[cutfeat] = cutval;
    int c1 = kd_divide(t, pts, vind, left, left + idx, llo, lhi);
    rlo[cutfeat] = cutval;
    int c2 = kd_divide(t, pts, vind, left + idx, right, rlo, rhi);
    kd_node nd;
    nd.left = nd.right = 0;
    nd.divfeat = cutfeat;
    nd.divlow = lhi[cutfeat];
    nd.divhigh = rlo[cutfeat];
    nd.child1 = c1;
    nd.child2 = c2;
    for (int i = 0; i < 3; i++) {
        lo[i] = llo[i] < rlo[i] ? llo[i] : rlo[i];
        hi[i] = lhi[i] > rhi[i] ? lhi[i] : rhi[i];
    }
    t->nodes[ni] = nd;
    return ni;
}

orc_kdtree *orc_kdtree_build(const orc_point_xyzi *pts, size_t m) {
    orc_kdtree *t = (orc_kdtree *) calloc(1, sizeof(orc_kdtree));
    t->m = m;
    if (m == 0) return t;
    float *raw = (float *) malloc(sizeof(float) * 3 * m); /* PCL copies xyz into a flat float array */
    t->vind = (int32_t *) malloc(sizeof(int32_t) * m);
    for (size_t i = 0; i < m; i++) {
        raw[3 * i] = pts[i].x;
        raw[3 * i + 1] = pts[i].y;
        raw[3 * i + 2] = pts[i].z;
        t->vind[i] = (int32_t) i;
    }
    for (int i = 0; i < 3; i++) t->bb_lo[i] = t->bb_hi[i] = raw[i];
    for (size_t k = 1; k < m; k++)
        for (int i = 0; i < 3; i++) {
            float v = raw[3 * k + i];
            if (v < t->bb_lo[i]) t->bb_lo[i] = v;
            if (v > t->bb_hi[i]) t->bb_hi[i] = v;
        }
    float lo[3], hi[3];
    memcpy(lo, t->bb_lo, sizeof(lo));
    memcpy(hi, t->bb_hi, sizeof(hi));
    kd_divide(t, raw, t->vind, 0, (int) m, lo, hi);
    t->data = (float *) malloc(sizeof(float) * 3 * m); /* reorder_ = true */
    for (size_t i = 0; i < m; i++) memcpy(&t->data[3 * i], &raw[3 * t->vind[i]], 3 * sizeof(float));
    free(raw);
    return t;
}
void orc_kdtree_destroy(orc_kdtree *t) {
    if (!t) return;
    free(t->data);
    free(t->vind);
    free(t->nodes);
    free(t);
}

typedef struct {
    int k, count;
    int32_t *idx;
    float *d2;
} kd_result;

static void kd_search(const orc_kdtree *t, int ni, const float q[3], float mindistsq, float dists[3], kd_result *res) {
    const kd_node *nd = &t->nodes[ni];
    if (nd->divfeat < 0) {
        float worst = (res->count == res->k) ? res->d2[res->k - 1] : FLT_MAX;
        for (int i = nd->left; i < nd->right; i++) {
            float d = l2_simple(q, &t->data[3 * i]);
            if (d <= worst) { /* <= so that (d2, index) ties resolve by index, not visit order */
                knn_insert(res->k, &res->count, res->idx, res->d2, d, t->vind[i]);
                worst = (res->count == res->k) ? res->d2[res->k - 1] : FLT_MAX;
            }
        }
        return;
    }
    int f = nd->divfeat;
    float val = q[f];
    float diff1 = val - nd->divlow;
    float diff2 = val - nd->divhigh;
    int best, other;
    float cut_dist;
    if ((diff1 + diff2) < 0) {
        best = nd->child1;
        other = nd->child2;
        cut_dist = (val - nd->divhigh) * (val - nd->divhigh);
    } else {
        best = nd->child2;
        other = nd->child1;
        cut_dist = (val - nd->divlow) * (val - nd->divlow);
    }
    kd_search(t, best, q, mindistsq, dists, res);
    float dst = dists[f];
    mindistsq = mindistsq + cut_dist - dst;
    dists[f] = cut_dist;
    float worst = (res->count == res->k) ? res->d2[res->k - 1] : FLT_MAX;
    if (mindistsq <= worst) kd_search(t, other, q, mindistsq, dists, res);
    dists[f] = dst;
}

int orc_kdtree_knn(const orc_kdtree *t, const float q[3], int k, int32_t *idx, float *d2) {
    if (t->m == 0) return 0;
    kd_result res = {k, 0, idx, d2};
    float dists[3] = {0, 0, 0};
    float distsq = 0;
    for (int i = 0; i < 3; i++) {
        if (q[i] < t->bb_lo[i]) {
            dists[i] = (q[i] - t->bb_lo[i]) * (q[i] - t->bb_lo[i]);
            distsq += dists[i];
        }
        if (q[i] > t->bb_hi[i]) {
            dists[i] = (q[i] - t->bb_hi[i]) * (q[i] - t->bb_hi[i]);
            distsq += dists[i];
        }
    }
    kd_search(t, 0, q, distsq, dists, &res);
    return res.count;
}

/* ------------------------------------------------------------------------------------------
 * Eigen::ColPivHouseholderQR<Matrix<double,5,3>>::compute + solve  (Eigen >= 3.3, ColPivHouseholderQR.h)
 * m is column-major 5x3 in q[col*5 + row].
 * ------------------------------------------------------------------------------------------ */
void orc_colpiv_qr_solve_5x3(const double A[15], const double b[5], double x[3]) {
    enum { ROWS = 5, COLS = 3, SIZE = 3 };
    double q[ROWS * COLS];
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++) q[c * ROWS + r] = A[r * COLS + c];
    double hcoef[SIZE];
    int perm[COLS] = {0, 1, 2};
    double norms_upd[COLS], norms_dir[COLS];
    for (int k = 0; k < COLS; k++) {
        double s = 0;
        for (int r = 0; r < ROWS; r++) s += q[k * ROWS + r] * q[k * ROWS + r];
        norms_dir[k] = sqrt(s);
        norms_upd[k] = norms_dir[k];
    }
    double maxn = norms_upd[0];
    for (int k = 1; k < COLS; k++)
        if (norms_upd[k] > maxn) maxn = norms_upd[k];
    double th = maxn * DBL_EPSILON / (double) ROWS;
    const double threshold_helper = th * th;
    const double norm_downdate_threshold = sqrt(DBL_EPSILON);
    int nonzero_pivots = SIZE;

    for (int k = 0; k < SIZE; k++) {
        int big = k;
        double bigv = norms_upd[k];
        for (int j = k + 1; j < COLS; j++)
            if (norms_upd[j] > bigv) {
                bigv = norms_upd[j];
                big = j;
            }
        double biggest_col_sq_norm = bigv * bigv;
        if (nonzero_pivots == SIZE && biggest_col_sq_norm < threshold_helper * (double) (ROWS - k)) nonzero_pivots = k;
        if (k != big) {
            for (int r = 0; r < ROWS; r++) {
                double tmp = q[k * ROWS + r];
                q[k * ROWS + r] = q[big * ROWS + r];
                q[big * ROWS + r] = tmp;
            }
            double tmp = norms_upd[k];
            norms_upd[k] = norms_upd[big];
            norms_upd[big] = tmp;
            tmp = norms_dir[k];
            norms_dir[k] = norms_dir[big];
            norms_dir[big] = tmp;
            int ti = perm[k];
            perm[k] = perm[big];
            perm[big] = ti;
        }
        /* makeHouseholderInPlace on q.col(k).tail(ROWS-k) */
        double *col = &q[k * ROWS + k];
        int len = ROWS - k;
        double tail_sq = 0;
        for (int r = 1; r < len; r++) tail_sq += col[r] * col[r];
        double c0 = col[0], beta, tau;
        if (tail_sq <= DBL_MIN) {
            tau = 0;
            beta = c0;
            for (int r = 1; r < len; r++) col[r] = 0;
        } else {
            beta = sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
            for (int r = 1; r < len; r++) col[r] = col[r] / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hcoef[k] = tau;
        col[0] = beta;
        /* bottomRightCorner(ROWS-k, COLS-k-1).applyHouseholderOnTheLeft(essential, tau) */
        if (tau != 0) {
            for (int j = k + 1; j < COLS; j++) {
                double *cj = &q[j * ROWS + k];
                double tmp = 0;
                for (int r = 1; r < len; r++) tmp += col[r] * cj[r];
                tmp += cj[0];
                cj[0] -= tau * tmp;
                for (int r = 1; r < len; r++) cj[r] -= (tau * col[r]) * tmp;
            }
        }
        /* LAPACK-style column-norm down-dating */
        for (int j = k + 1; j < COLS; j++) {
            if (norms_upd[j] != 0) {
                double temp = fabs(q[j * ROWS + k]) / norms_upd[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0 ? 0 : temp;
                double ratio = norms_upd[j] / norms_dir[j];
                double temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    double s = 0;
                    for (int r = k + 1; r < ROWS; r++) s += q[j * ROWS + r] * q[j * ROWS + r];
                    norms_dir[j] = sqrt(s);
                    norms_upd[j] = norms_dir[j];
                } else {
                    norms_upd[j] *= sqrt(temp);
                }
            }
        }
    }

    /* solve: c = Q^T b (H_0 first), back-substitute R, un-permute */
    double c[ROWS];
    for (int r = 0; r < ROWS; r++) c[r] = b[r];
    for (int k = 0; k < nonzero_pivots; k++) {
        double tau = hcoef[k];
        int len = ROWS - k;
        const double *ess = &q[k * ROWS + k];
        if (tau != 0) {
            double tmp = 0;
            for (int r = 1; r < len; r++) tmp += ess[r] * c[k + r];
            tmp += c[k];
            c[k] -= tau * tmp;
            for (int r = 1; r < len; r++) c[k + r] -= (tau * ess[r]) * tmp;
        }
    }
    for (int i = nonzero_pivots - 1; i >= 0; i--) { /* column-oriented back substitution */
        c[i] /= q[i * ROWS + i];
        for (int r = 0; r < i; r++) c[r] -= c[i] * q[i * ROWS + r];
    }
    for (int i = 0; i < COLS; i++) x[i] = 0;
    for (int i = 0; i < nonzero_pivots; i++) x[perm[i]] = c[i];
}

/* pointcloud.cc:73-105  PointCloudCommon::planeEstimation (no call site in the reference) */
int orc_plane_estimation(const orc_point_xyzi *points, size_t n, double threshold, double unit_normal[3],
                         double *norm_inverse, double distance[5], int *n_distance) {
    *n_distance = 0;
    if (n < 5) return 0;
    double A[15], b[5], nv[3];
    for (int k = 0; k < 5; k++) {
        A[3 * k] = (double) points[k].x;
        A[3 * k + 1] = (double) points[k].y;
        A[3 * k + 2] = (double) points[k].z;
        b[k] = -1.0;
    }
    orc_colpiv_qr_solve_5x3(A, b, nv);
    *norm_inverse = 1.0 / v3d_norm(nv);
    unit_normal[0] = nv[0] * *norm_inverse;
    unit_normal[1] = nv[1] * *norm_inverse;
    unit_normal[2] = nv[2] * *norm_inverse;
    for (int k = 0; k < 5; k++) {
        double dis = v3d_dot(unit_normal, &A[3 * k]) + *norm_inverse;
        distance[(*n_distance)++] = dis;
        if (fabs(dis) > threshold) return 0;
    }
    return 1;
}

/* symmetric 3x3 eigen-decomposition (cyclic Jacobi); stands in for Eigen::EigenSolver on the covariance at
 * pointcloud.cc:132-134.  w: eigenvalues, v: eigenvectors as columns (row-major). */
static void sym_eigen3(const double a_in[9], double w[3], double v[9]) {
    double a[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a[i][j] = a_in[3 * i + j];
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; k++) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = cs * akp - sn * akq;
                    a[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; k++) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = cs * apk - sn * aqk;
                    a[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = V[k][p], vkq = V[k][q];
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

/* pointcloud.cc:111-190  PointCloudCommon::pointToPlaneStd (no call site in the reference) */
double orc_point_to_plane_std(const orc_point_xyzi *nearest, size_t n, const orc_point_xyzi *target, double point_std) {
    const double pvv = point_std * point_std; /* :113-114 */
    const double inv = 1.0 / (double) n;      /* :117 */
    double cov[9] = {0}, center[3] = {0};
    for (size_t i = 0; i < n; i++) { /* :120-126 */
        double pv[3] = {nearest[i].x, nearest[i].y, nearest[i].z};
        for (int r = 0; r < 3; r++) {
            center[r] += pv[r];
            for (int c = 0; c < 3; c++) cov[3 * r + c] += pv[r] * pv[c];
        }
    }
    for (int r = 0; r < 3; r++) center[r] *= inv; /* :127 */
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) cov[3 * r + c] = cov[3 * r + c] * inv - center[r] * center[c]; /* :128 */
    double evals[3], evecs[9];
    sym_eigen3(cov, evals, evecs); /* :131-136 */
    int emin = 0;                  /* :138-140 */
    for (int m = 1; m < 3; m++)
        if (evals[m] < evals[emin]) emin = m;
    double nrm[3] = {evecs[emin], evecs[3 + emin], evecs[6 + emin]}; /* :151 */
    double plane_cov[6][6];
    memset(plane_cov, 0, sizeof(plane_cov));
    for (size_t i = 0; i < n; i++) { /* :160-176 */
        double d[3] = {nearest[i].x - center[0], nearest[i].y - center[1], nearest[i].z - center[2]};
        double F[3][3];
        memset(F, 0, sizeof(F));
        for (int m = 0; m < 3; m++) {
            if (m == emin) continue;
            double em[3] = {evecs[m], evecs[3 + m], evecs[6 + m]};
            double scale = inv / (evals[emin] - evals[m]);
            double de = d[0] * em[0] + d[1] * em[1] + d[2] * em[2], dn = d[0] * nrm[0] + d[1] * nrm[1] + d[2] * nrm[2];
            for (int c = 0; c < 3; c++) F[m][c] = scale * (de * nrm[c] + dn * em[c]);
        }
        double J[6][3];
        memset(J, 0, sizeof(J));
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                for (int m = 0; m < 3; m++) J[r][c] += evecs[3 * r + m] * F[m][c];
        for (int r = 0; r < 3; r++) J[3 + r][r] = inv;
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) {
                double sum = 0;
                for (int k = 0; k < 3; k++) sum += J[r][k] * J[c][k];
                plane_cov[r][c] += pvv * sum;
            }
    }
    double jac[6] = {target->x - center[0], target->y - center[1], target->z - center[2], -nrm[0], -nrm[1], -nrm[2]}; /* :179-181 */
    double var = 0;
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) var += jac[r] * plane_cov[r][c] * jac[c]; /* :184 */
    var += pvv * (nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);   /* :186 */
    return sqrt(var);
}

/* ------------------------------------------------------------------------------------------
 * tracking.cc:564-634  Tracking::depthAssociationPlaneFitting
 * ------------------------------------------------------------------------------------------ */
int orc_plane_fitting(const orc_point_xyzi *us, size_t m, const orc_kdtree *tree, orc_point_xyzi *query,
                      int32_t *knn_idx_out, float *knn_d2_out, int *flags_out) {
    int32_t k_indices[5] = {-1, -1, -1, -1, -1};
    float k_sqr[5] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX};
    int type = ORC_MAPPOINT_NONE;
    int flags = 0;
    const float qv[3] = {query->x, query->y, query->z};

    int found = tree ? orc_kdtree_knn(tree, qv, 5, k_indices, k_sqr) : orc_brute_knn(us, m, qv, 5, k_indices, k_sqr); /* :571 */
    if (knn_idx_out) memcpy(knn_idx_out, k_indices, sizeof(k_indices));
    if (knn_d2_out) memcpy(knn_d2_out, k_sqr, sizeof(k_sqr));

    if (found < 5) { /* reference reads k_sqr_distances[4] out of bounds (UB); the build defines NONE */
        flags |= 1;
        if (flags_out) *flags_out = flags;
        return type;
    }
    if ((double) k_sqr[4] > orc_const_nearest_sq_threshold()) { /* :573-575 */
        flags |= 2;
        if (flags_out) *flags_out = flags;
        return type;
    }

    double P[5][3]; /* :577-581  pt.xyz.cast<double>() * pt.intensity */
    for (int k = 0; k < 5; k++) {
        const orc_point_xyzi *pt = &us[k_indices[k]];
        double r = (double) pt->intensity;
        P[k][0] = (double) pt->x * r;
        P[k][1] = (double) pt->y * r;
        P[k][2] = (double) pt->z * r;
    }

    double ma[15], mb[5], mx[3]; /* :584-589 */
    for (int k = 0; k < 5; k++) {
        ma[3 * k] = P[k][0];
        ma[3 * k + 1] = P[k][1];
        ma[3 * k + 2] = P[k][2];
        mb[k] = -1.0;
    }
    orc_colpiv_qr_solve_5x3(ma, mb, mx);

    double norm_inverse = 1.0 / v3d_norm(mx); /* :592 */
    {                                          /* :593 mx.normalize() */
        double z = v3d_dot(mx, mx);
        if (z > 0) {
            double s = sqrt(z);
            mx[0] /= s;
            mx[1] /= s;
            mx[2] /= s;
        }
    }

    int isvalid = 1; /* :596-605 */
    for (int k = 0; k < 5; k++) {
        double sm = fabs(v3d_dot(mx, P[k]) + norm_inverse);
        if (sm > 0.1) {
            isvalid = 0;
            flags |= 4;
            break;
        }
    }

    float depth_diff = fabsf(us[k_indices[0]].intensity - us[k_indices[4]].intensity); /* :607-608 */
    if ((double) depth_diff > 0.5) {                                                   /* :609-611 */
        isvalid = 0;
        flags |= 8;
    }
    if (isvalid) { /* :612-631 */
        float vr[3] = {(float) P[4][0], (float) P[4][1], (float) P[4][2]};
        float vi[3] = {query->x, query->y, query->z};
        float vn[3] = {(float) mx[0], (float) mx[1], (float) mx[2]};

        float t = v3f_dot(vn, vr) / v3f_dot(vn, vi); /* :618 */

        if (isnan(t) || ((double) t < NEAREST_DEPTH) || ((double) t > FARTHEST_DEPTH)) { /* :621-623 */
            flags |= 16;
            if (flags_out) *flags_out = flags;
            return type;
        }
        query->x *= t; /* :626 */
        query->y *= t;
        query->z *= t;
        query->intensity = query->z; /* :629 */
        type = ORC_MAPPOINT_DEPTH_ASSOCIATED;
    }
    if (flags_out) *flags_out = flags;
    return type;
}

void orc_associate(const orc_point_xyzi *us, size_t m, const orc_kdtree *tree, const float *queries, size_t f,
                   int8_t *type, float *xyz_depth, int32_t *knn_idx, float *knn_d2, int32_t *flags) {
    for (size_t k = 0; k < f; k++) { /* serial like tracking.cc:399-413 / 512-525 */
        orc_point_xyzi q = {queries[3 * k], queries[3 * k + 1], queries[3 * k + 2], 1.0f, 0, {0, 0, 0}};
        int fl = 0;
        int ty = orc_plane_fitting(us, m, tree, &q, knn_idx ? &knn_idx[5 * k] : NULL, knn_d2 ? &knn_d2[5 * k] : NULL, &fl);
        type[k] = (int8_t) ty;
        xyz_depth[4 * k] = q.x;
        xyz_depth[4 * k + 1] = q.y;
        xyz_depth[4 * k + 2] = q.z;
        xyz_depth[4 * k + 3] = q.intensity;
        if (flags) flags[k] = fl;
    }
}

/* ------------------------------------------------------------------------------------------
 * camera.cc:123-127, 133-135, 149-151; visual_mappoint.cc:90-92
 * ------------------------------------------------------------------------------------------ */
void orc_pixel2cam(const double intr[5], float u, float v, double pc[3]) {
    double y = ((double) v - intr[3]) / intr[1];
    double x = ((double) u - intr[2] - intr[4] * y) / intr[0];
    pc[0] = x;
    pc[1] = y;
    pc[2] = 1.0;
}
void orc_pixel2unitcam(const double intr[5], float u, float v, double puc[3]) {
    orc_pixel2cam(intr, u, v, puc);
    double z = v3d_dot(puc, puc);
    if (z > 0) {
        double s = sqrt(z);
        puc[0] /= s;
        puc[1] /= s;
        puc[2] /= s;
    }
}
void orc_cam2world(const double pc[3], const orc_pose *pose, double pw[3]) {
    double r[3];
    m3_mul_v(pose->R, pc, r);
    pw[0] = r[0] + pose->t[0];
    pw[1] = r[1] + pose->t[1];
    pw[2] = r[2] + pose->t[2];
}
/* cv::undistortPoints(pts, pts, K, D, Mat(), K) as Camera::undistortPoints calls it (camera.cc:72-74; OpenCV >= 3.2,
 * README.md:40): OpenCV's cvUndistortPointsInternal with the default criteria (5 fixed-point iterations, no epsilon test),
 * R = identity and P = K, restated operation for operation.  D = (k1, k2, p1, p2, k3); OpenCV normalises with fx, fy, cx, cy
 * only (the skew K[0][1] enters when the point is projected back with P = K).  The input pixel is f32, the arithmetic f64,
 * the result is rounded to f32 (Point2f).  Checked bit for bit against this image's cv2.undistortPoints in
 * tests/test_oracle.py. */
void orc_undistort_points(const double intr[5], const double dist[5], const float *px_in, size_t n, float *px_out) {
    const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3], skew = intr[4];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k0 = dist[0], k1 = dist[1], k2 = dist[2], k3 = dist[3], k4 = dist[4];
    for (size_t i = 0; i < n; i++) {
        const double u = (double) px_in[2 * i], v = (double) px_in[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
            if (icdist < 0) {
                x = (u - cx) * ifx;
                y = (v - cy) * ify;
                break;
            }
            const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2;
            const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y + 0 * r2 + 0 * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = fx * x + skew * y + cx;
        const double yy = 0 * x + fy * y + cy;
        const double ww = 1. / (0 * x + 0 * y + 1);
        px_out[2 * i] = (float) (xx * ww);
        px_out[2 * i + 1] = (float) (yy * ww);
    }
}

int orc_is_good_depth(double depth, double scale) { return (depth > NEAREST_DEPTH) && (depth < FARTHEST_DEPTH * scale); }

/* ------------------------------------------------------------------------------------------
 * factors/depth_factor.h:38-42 as added at optimizer.cc:659-665 (sigma = ASSOCIATE_DEPTH_STD = 0.1, ceres::HuberLoss(1.0)),
 * and the chi-square test of optimizer.cc:752-783.  Ceres Solver (>= 2.1.0, README.md:48) is not vendored: the residual is
 * the functor's own expression, the Jacobian is its derivative (what ceres::AutoDiffCostFunction evaluates with Jets:
 * d/d rho of -1/rho is 1/rho^2), the loss is ceres::HuberLoss::Evaluate (loss_function.cc: rho(s) = s for s <= a^2, else
 * 2 a sqrt(s) - a^2) and a residual block's cost is 0.5 rho(|r|^2).
 * ------------------------------------------------------------------------------------------ */
void orc_depth_factor(double depth, double invdepth, double sigma, double huber_a, double *residual, double *jacobian, double *cost) {
    const double r = (depth - 1.0 / invdepth) / sigma;
    const double s = r * r;
    double rho = s;
    if (huber_a > 0) {
        const double b = huber_a * huber_a;
        if (s > b) rho = 2.0 * huber_a * sqrt(s) - b;
    }
    *residual = r;
    *jacobian = (1.0 / (invdepth * invdepth)) / sigma;
    *cost = 0.5 * rho;
}
int orc_chi2_outlier(const double *residual, size_t dim, double threshold, double *cost_out) {
    double c = 0;
    for (size_t i = 0; i < dim; i++) c += residual[i] * residual[i];
    if (cost_out) *cost_out = c;
    return c > threshold;
}

/* ------------------------------------------------------------------------------------------
 * whole keyframe with the reference's structure (bench CPU legs)
 * ------------------------------------------------------------------------------------------ */
size_t orc_keyframe(const orc_point_xyzit *buf, size_t n_buf, double stamp, double stack_interval, const orc_ins *ins,
                    size_t n_ins, const orc_pose *pose_b_c, const orc_pose *pose_c_l, double td_b_l,
                    const float *queries, size_t f, int libm_mode, int nthreads, int8_t *type, float *xyz_depth,
                    size_t *n_projected_out) {
    orc_point_xyzi *cloud = (orc_point_xyzi *) malloc(sizeof(orc_point_xyzi) * (n_buf ? n_buf : 1));
    size_t first = 0, n_out = 0;
    int ok = orc_stack_deskew(buf, n_buf, stamp, stack_interval, ins, n_ins, pose_b_c, pose_c_l, td_b_l, &first, &n_out,
                              cloud, NULL, NULL, nthreads);
    if (n_projected_out) *n_projected_out = n_out;
    size_t m = 0;
    for (size_t k = 0; k < f; k++) {
        type[k] = ORC_MAPPOINT_NONE;
        xyz_depth[4 * k] = queries[3 * k];
        xyz_depth[4 * k + 1] = queries[3 * k + 1];
        xyz_depth[4 * k + 2] = queries[3 * k + 2];
        xyz_depth[4 * k + 3] = 0;
    }
    if (ok > 0) { /* tracking.cc:162 !pointcloud.empty() */
        int32_t *winner = (int32_t *) malloc(sizeof(int32_t) * ORC_IMAGE_CELLS);
        orc_point_xyzi *us = (orc_point_xyzi *) malloc(sizeof(orc_point_xyzi) * ORC_IMAGE_CELLS);
        m = orc_range_image(cloud, n_out, libm_mode, NULL, winner, us, NULL);
        if (m > 0) { /* tracking.cc:165 */
            orc_kdtree *tree = orc_kdtree_build(us, m);
            orc_associate(us, m, tree, queries, f, type, xyz_depth, NULL, NULL, NULL);
            orc_kdtree_destroy(tree);
        }
        free(winner);
        free(us);
    }
    free(cloud);
    return m;
}

/* ------------------------------------------------------------------------------------------
 * ROS/lidar/lidar_converter.cc (driver records -> PointXYZIT), le_vins/common/gpstime.h
 * ------------------------------------------------------------------------------------------ */
double orc_unix2gps_sow(double unixs) { /* gpstime.h:36-41, GPS_LEAP_SECOND = 18 (:27) */
    double seconds = unixs + 18 - 315964800;
    int week = (int) floor(seconds / 604800);
    return seconds - week * 604800; /* int * int, then converted: like the reference */
}

static uint32_t rd_u32(const unsigned char *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }
static float rd_f32(const unsigned char *p) {
    uint32_t u = rd_u32(p);
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static double rd_f64(const unsigned char *p) {
    uint64_t u = (uint64_t) rd_u32(p) | ((uint64_t) rd_u32(p + 4) << 32);
    double d;
    memcpy(&d, &u, 8);
    return d;
}

size_t orc_convert_raw(const orc_raw_format *fmt, const void *records, size_t n, uint64_t timebase_ns, double stamp,
                       int to_gps_time, int scan_line, double nearest, double farthest, orc_point_xyzit *out, double *start,
                       double *end) {
    const unsigned char *base = (const unsigned char *) records;
    const double near2 = nearest * nearest, far2 = farthest * farthest; /* lidar_converter.cc:31-32 */
    double st = DBL_MAX, en = 0;                                          /* :50-51 */
    size_t m = 0;
    if (fmt->kind == 2 && fmt->time_kind != 3 && to_gps_time) stamp = orc_unix2gps_sow(stamp); /* :92-98, :141-147 */
    for (size_t k = 0; k < n; k++) {
        const unsigned char *r = base + k * (size_t) fmt->point_step;
        orc_point_xyzit p;
        memset(&p, 0, sizeof(p));
        if (fmt->kind == 1) {
            const unsigned line = r[fmt->off_ring], tag = r[fmt->off_tag];
            if (!((int) line < scan_line && ((tag & 0x30) == 0x00 || (tag & 0x30) == 0x10))) continue; /* :55 */
            p.x = rd_f32(r + fmt->off_x);
            p.y = rd_f32(r + fmt->off_y);
            p.z = rd_f32(r + fmt->off_z);
            p.intensity = (float) r[fmt->off_intensity];                                       /* :244 */
            p.time = (double) (timebase_ns + (uint64_t) rd_u32(r + fmt->off_time)) * 1.0e-9;  /* :245 */
            if (to_gps_time) p.time = orc_unix2gps_sow(p.time);
        } else {
            /* point_counts[ring] % point_filter_num_ == 0 with point_filter_num_ == 1: always (:108) */
            p.x = rd_f32(r + fmt->off_x);
            p.y = rd_f32(r + fmt->off_y);
            p.z = rd_f32(r + fmt->off_z);
            p.intensity = rd_f32(r + fmt->off_intensity);
            if (fmt->time_kind == 1)
                p.time = stamp + rd_f32(r + fmt->off_time); /* :113 */
            else if (fmt->time_kind == 2)
                p.time = stamp + rd_u32(r + fmt->off_time) * 1.0e-9; /* :162 */
            else {
                p.time = rd_f64(r + fmt->off_time); /* :213-221 */
                if (to_gps_time) p.time = orc_unix2gps_sow(p.time);
            }
        }
        if (p.time < st) st = p.time; /* :58-63 */
        if (p.time > en) en = p.time;
        const double sq = (double) v3f_sqnorm(p.x, p.y, p.z); /* :66 Vector3f::squaredNorm() */
        if (sq > near2 && sq < far2) out[m++] = p;               /* :67-69 */
    }
    *start = st;
    *end = en;
    return m;
}
