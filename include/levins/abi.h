/*
 * levins/abi.h -- C ABI of liblevins_b200.so: the B200-native LiDAR depth-enhancement path of LE-VINS.
 *
 * The reference (i2Nav-WHU/LE-VINS) has no plugin / FFI boundary for this path: it is plain C++ member
 * and static functions inside the estimator.  This header is therefore the boundary a
 * maintainer binds instead; every entry point names the reference code it replaces (file:line, relative
 * to the reference checkout).  Plain pointers and sizes only; no C++/torch types.  All host pointers are
 * caller-owned.  The library has NO CPU fallback: every compute entry point runs hand-written sm_100a
 * kernels and fails with LEVINS_ERR_CUDA when no such device is present.
 *
 * Threading: a context is bound to one GPU and is not thread-safe, except levins_append_scan(), which may
 * be called from another thread (the reference appends from the ROS callback thread under a mutex,
 * le_vins.cc:149-163).  As in the reference, whose stack loop holds that mutex (le_vins.cc:392), an append that arrives
 * while levins_keyframe / levins_stack_deskew is running waits until that call has returned (the context's mutex is
 * held across the call's device synchronisation: ~0.08 ms for a keyframe).
 */
#ifndef LEVINS_ABI_H
#define LEVINS_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LEVINS_ABI_VERSION 2

/* ---- status codes (the reference has bool/int returns and no error codes) ---- */
#define LEVINS_OK 0
#define LEVINS_EMPTY 1            /* the reference's `return false` / empty cloud (le_vins.cc:426-429, tracking.cc:356-358) */
#define LEVINS_ERR_ARG (-1)
#define LEVINS_ERR_CUDA (-2)
#define LEVINS_ERR_CAPACITY (-3)
#define LEVINS_ERR_STATE (-4)

/* ---- record layouts (byte-identical to the reference's) ---- */

/* PointXYZIT, le_vins/lidar/lidar.h:28-33: 32 B, 16-aligned, time = GPS seconds of week (f64). */
typedef struct levins_point_xyzit {
    float x, y, z, pad0;
    float intensity, pad1;
    double time;
} levins_point_xyzit;

/* PointType = pcl::PointXYZI, le_vins/lidar/lidar.h:89: 32 B; w (data[3]) = 1.0f. */
typedef struct levins_point_xyzi {
    float x, y, z, w;
    float intensity, pad[3];
} levins_point_xyzi;

/* Pose, le_vins/common/types.h:51-54: Matrix3d R (column-major) + Vector3d t. */
typedef struct levins_pose {
    double R[9];
    double t[3];
} levins_pose;

/* One element of ins_window_ (std::deque<std::pair<IMU, IntegrationState>>, le_vins.h):
 * `time` = pair.first.time (what MISC::getInsWindowIndex compares, misc.cc:28,40-41),
 * `state_time`, `p`, `q` = pair.second.{time,p,q} (what statePoseInterpolation reads, misc.cc:84-94).
 * q = (x, y, z, w), Eigen coefficient order. */
typedef struct levins_ins_sample {
    double time;
    double state_time;
    double p[3];
    double q[4];
} levins_ins_sample;

/* MapPointType, le_vins/visual/visual_mappoint.h:38-42 */
#define LEVINS_MAPPOINT_NONE (-1)
#define LEVINS_MAPPOINT_TRIANGULATED 0
#define LEVINS_MAPPOINT_DEPTH_ASSOCIATED 1

/* range image geometry: compile-time constants in the reference (tracking.h:119-126) and here */
#define LEVINS_IMAGE_SIZE 300
#define LEVINS_IMAGE_CELLS (LEVINS_IMAGE_SIZE * LEVINS_IMAGE_SIZE)

/* decision-path bits reported per feature (diagnostic, same meaning as the oracle's) */
#define LEVINS_FLAG_FEW_POINTS 1    /* fewer than 5 neighbours exist (reference: UB, build: NONE) */
#define LEVINS_FLAG_TOO_FAR 2       /* tracking.cc:573-575 */
#define LEVINS_FLAG_RESIDUAL 4      /* tracking.cc:596-605 */
#define LEVINS_FLAG_RANGE_SPREAD 8  /* tracking.cc:607-611 */
#define LEVINS_FLAG_BAD_T 16        /* tracking.cc:621-623 */

/* All association thresholds are compile-time constants in the reference; they are run-time here with
 * the reference's values as defaults (levins_config_default). */
typedef struct levins_config {
    uint32_t struct_size;            /* sizeof(levins_config), for forward compatibility */
    int32_t libm_mode;               /* 0: unqualified atan2/sqrt resolve to the float overloads (default), 1: double (SURVEY 8a note 3) */
    double stack_interval;           /* LIDAR_DEPTH_POINT_STACK_INTERVAL = 0.5 s, le_vins.h:130 */
    double reserved_buffer_length;   /* RESERVED_BUFFER_LENGTH = 10 s; eviction at length + 2*interval, le_vins.cc:156 */
    double nearest_depth;            /* VisualMapPoint::NEAREST_DEPTH = 1, visual_mappoint.h:50 */
    double farthest_depth;           /* VisualMapPoint::FARTHEST_DEPTH = 200, visual_mappoint.h:51 */
    double plane_residual_threshold; /* 0.1 m, tracking.cc:600 */
    double range_spread_threshold;   /* 0.5 m, tracking.cc:609 */
    int32_t nearest_distance_counts; /* ASSOCIATE_NEAREST_DISTANCE_COUNTS = 3, tracking.h:121 */
    int32_t deskew_precision;        /* 0 (default): the reference's f64 pose algebra, deskewed coordinates bit-identical to the
                                      * reference; 1: f32 algebra (f64 time differences), coordinates within ~1e-6 relative
                                      * (bar: 1e-4), cell membership of points next to a cell boundary may differ */
    uint64_t max_points;             /* device ring-buffer capacity in points (streaming API) */
    uint64_t max_ins;                /* INS samples per keyframe (reference: MAXIMUM_INS_NUMBER 1000 + RESERVED 400) */
    uint64_t max_features;           /* feature queries per keyframe */
} levins_config;

typedef struct levins_ctx levins_ctx;

/* ------------------------------------------------------------------------------------------------
 * lifecycle
 * ------------------------------------------------------------------------------------------------ */
int levins_abi_version(void);
void levins_config_default(levins_config *cfg);
/* device = CUDA ordinal.  Fails with LEVINS_ERR_CUDA when the device is missing or is not sm_100. */
int levins_create(int device, const levins_config *cfg, levins_ctx **out);
void levins_destroy(levins_ctx *ctx);
const char *levins_last_error(const levins_ctx *ctx);
const char *levins_status_string(int status);
/* Run on an externally owned cudaStream_t (e.g. torch's current stream) instead of the context's own. */
int levins_set_stream(levins_ctx *ctx, void *cuda_stream);
int levins_synchronize(levins_ctx *ctx);

/* pose_b_c_, pose_c_l_, td_b_l_ of class VINS (le_vins.cc:83-99) */
int levins_set_extrinsics(levins_ctx *ctx, const levins_pose *pose_b_c, const levins_pose *pose_c_l, double td_b_l);

/* ------------------------------------------------------------------------------------------------
 * single-shot pose lookups on the host (no GPU involved): what the estimator itself asks MISC for once per frame
 * (le_vins.cc:364).  The per-point use of the same functions inside the deskew loop (le_vins.cc:455) is what the kernels
 * replace; these four share their arithmetic (one source, compiled for host and device).
 * ------------------------------------------------------------------------------------------------ */
/* MISC::getInsWindowIndex (misc.cc:25-60): i with time[i-1] <= t < time[i]; 0 when the window is empty, t < time[0] or
 * t >= time[n-1], or when no bracketing pair exists. */
size_t levins_ins_window_index(const levins_ins_sample *window, size_t n, double time);
/* MISC::statePoseInterpolation (misc.cc:80-95): out->{state_time, p, q} between two samples; out->time = midtime. */
void levins_state_interpolate(const levins_ins_sample *s0, const levins_ins_sample *s1, double midtime, levins_ins_sample *out);
/* MISC::stateToPoseWithExtrinsic (misc.cc:97-103) */
void levins_state_to_pose(const levins_ins_sample *state, const levins_pose *pose_ext, levins_pose *out);
/* MISC::getPoseFromInsWindow (misc.cc:62-78): returns 1 when interpolated, 0 when the newest state was used. */
int levins_pose_at(const levins_ins_sample *window, size_t n, const levins_pose *pose_ext, double time, levins_pose *out);

/* ------------------------------------------------------------------------------------------------
 * streaming API -- one keyframe at a time, host buffers in and out
 * ------------------------------------------------------------------------------------------------ */

/* VINS::addNewPointCloud (le_vins.cc:149-163): append one scan to the device-resident lidar_point_buffer_
 * and evict from the front while |t_back - t_front| > reserved_buffer_length + 2*stack_interval.
 * Points must be time-sorted within and across scans (the reference assumes it, le_vins.cc:431-437). */
int levins_append_scan(levins_ctx *ctx, const levins_point_xyzit *points, size_t n);

/* Page-locked host memory for scan buffers (optional).  A scan that lives in such memory is copied to the device
 * asynchronously at PCIe speed (levins_append_scan returns after ~10 us); a scan in ordinary pageable memory -- e.g. the
 * pcl::PointCloud the ROS converter fills (ROS/lidar/lidar_converter.cc:35-256) -- is staged by the driver first and
 * costs ~80 us per 24 000 points.  Typical use: allocate two or three scan-sized buffers once and let the converter
 * write PointXYZIT records into them in turn. */
int levins_alloc_host(size_t bytes, void **out);
/* levins_append_scan returns before the device has read the caller's buffer when that buffer is page-locked (the copy is
 * ordered on the context's stream behind whatever is queued there).  A buffer may be overwritten again once
 * levins_append_wait() has returned (it waits for the copies of the most recent levins_append_scan / levins_append_raw), or
 * after any call that
 * synchronises the context (levins_keyframe, levins_stack_deskew, levins_synchronize).  Rotating two or three buffers and
 * waiting before re-filling the oldest keeps the converter and the copy overlapped.  Scans in pageable memory are staged by
 * the driver before the call returns and need no wait. */
int levins_append_wait(levins_ctx *ctx);
int levins_free_host(void *p);

/* Driver records straight into the ring: LidarConverter::{livox,velodyne,ouster,hesai}PointCloudConvertion
 * (ROS/lidar/lidar_converter.cc:35-256) + VINS::addNewPointCloud in one call.  The records are copied to the device as
 * they are (Livox: 19-20 B instead of 32 B per point) and filtered, converted and compacted there, in record order:
 *   kind 1  Livox CustomPoint: u32 offset_time, f32 x y z, u8 reflectivity, u8 tag, u8 line.  Kept iff line < scan_line and
 *           (tag & 0x30) is 0x00 or 0x10 (:55); time = (double)(timebase_ns + offset_time) * 1e-9 (:245).
 *   kind 2  PointCloud2-style record: f32 x y z intensity, ring, and a time field: time_kind 1 = f32 seconds added to
 *           `stamp` (Velodyne :113), 2 = u32 nanoseconds added to `stamp` (Ouster :162), 3 = f64 unix seconds (Hesai :213).
 * to_gps_time converts unix seconds to GPS seconds of week (le_vins/common/gpstime.h:36-41).  A record survives iff
 * nearest^2 < |p|^2 < farthest^2 (:66-67).  point_filter_num_ is the constant 1 in the reference (lidar_converter.h:61).
 * Outputs: the number of points appended and the converter's start / end times (over every considered record, :58-63).
 * Reporting them needs a device->host read, i.e. a synchronisation; with all three output pointers NULL nothing is read back
 * and the call only enqueues (one copy, two launches; levins_append_wait tells when `records` has been read).
 * Fields may sit at any byte offset and the stride may be odd (the serialized 19-byte Livox record). */
typedef struct levins_raw_format {
    uint32_t kind, point_step;
    uint32_t off_x, off_y, off_z, off_intensity, off_time, off_ring, off_tag;
    uint32_t time_kind;  /* kind 2 only */
    uint32_t ring_bytes; /* 1 or 2 (unused: every record is considered) */
    uint32_t pad;
} levins_raw_format;
int levins_append_raw(levins_ctx *ctx, const levins_raw_format *fmt, const void *records, size_t n, uint64_t timebase_ns,
                      double stamp, int to_gps_time, int scan_line, double nearest, double farthest, size_t *n_appended,
                      double *t_start, double *t_end);
/* current deque size and its front/back times (synchronises) */
int levins_buffer_info(levins_ctx *ctx, size_t *n_points, double *t_front, double *t_back);
int levins_buffer_clear(levins_ctx *ctx);

/* VINS::pointcloudStackForVisualTracking (le_vins.cc:392-470) + MISC::getPoseFromInsWindow (misc.cc:25-103)
 * + PointCloudCommon::relativeProjectionToCamera (pointcloud.cc:36-48,65-71):
 * pops points older than stamp - stack_interval, takes the points up to `stamp`, range-gates and deskews
 * each with the INS pose interpolated at time + td_b_l into the camera frame of the keyframe at `stamp`.
 * out_cloud (nullable): receives n_out points (gated points are default points (0,0,0,1;0), as after
 * PointCloud::resize, le_vins.cc:438); out_capacity in points.  The deskewed cloud also stays resident
 * on the device for levins_build_range_image(ctx, NULL, ...).
 * Returns LEVINS_OK, LEVINS_EMPTY (reference returns false), or an error. */
int levins_stack_deskew(levins_ctx *ctx, double stamp, const levins_ins_sample *ins, size_t n_ins,
                        levins_point_xyzi *out_cloud, size_t out_capacity, size_t *n_out);

/* Tracking::depthAssociationProjection (tracking.cc:298-368): 300x300 spherical range image that keeps the
 * nearest point per 0.3-degree cell, row-major compaction to the unit sphere (xyz = unit vector,
 * intensity = range).  The KD-tree build of tracking.cc:361-362 has no counterpart: the cell grid is the index.
 * cloud == NULL: use the resident result of the last levins_stack_deskew.
 * unit_sphere_out (nullable, capacity unit_capacity), cell_winner_out (nullable, 90000 entries: index into the
 * cloud of the point kept per cell or -1), m_out = number of non-empty cells.
 * Returns LEVINS_OK, LEVINS_EMPTY when M == 0 (reference returns false), or an error. */
int levins_build_range_image(levins_ctx *ctx, const levins_point_xyzi *cloud, size_t n,
                             levins_point_xyzi *unit_sphere_out, size_t unit_capacity, int32_t *cell_winner_out,
                             size_t *m_out);

/* Tracking::depthAssociationPlaneFitting (tracking.cc:564-634) for F query rays against the range image built
 * by the last levins_build_range_image / levins_keyframe (the per-feature loops of tracking.cc:399-413 and
 * 512-525).  query_xyz: F x 3 f32 unit vectors (Camera::pixel2unitcam cast to float, tracking.cc:387-395).
 * Outputs (each nullable): type[F] MapPointType; xyz_depth[F*4] = camera-frame x,y,z and depth (= z), the
 * query unchanged with depth 0 when type is NONE; knn_idx[F*5] indices into the compacted unit-sphere cloud,
 * ascending (d2, index), -1 where fewer exist inside the search window; knn_d2[F*5]; flags[F] LEVINS_FLAG_*.
 * Neighbours farther than the acceptance threshold (tracking.cc:573) can never influence a result and are only
 * guaranteed to be reported when they lie inside the 9 x 13-cell search window. */
int levins_associate(levins_ctx *ctx, const float *query_xyz, size_t n_features, int8_t *type, float *xyz_depth,
                     int32_t *knn_idx, float *knn_d2, int32_t *flags);

/* Fused keyframe: levins_stack_deskew + levins_build_range_image + levins_associate with no intermediate
 * host round trip (the deskewed cloud is never written; one device->host copy of the results).
 * n_projected_out / m_out nullable. */
int levins_keyframe(levins_ctx *ctx, double stamp, const levins_ins_sample *ins, size_t n_ins, const float *query_xyz,
                    size_t n_features, int8_t *type, float *xyz_depth, size_t *n_projected_out, size_t *m_out);

/* The feature side of the two association drivers on the device: tracked pixels in, accepted map points
 * out.  Per feature: Camera::undistortPoints (camera.cc:72-74: cv::undistortPoints(pts, pts, K, D, Mat(), K), OpenCV's
 * 5-iteration fixed point; skipped when `undistort` is 0 because the caller's pixels are undistorted already) ->
 * Camera::pixel2unitcam -> cast to float (tracking.cc:381-395 / 489-504) -> depthAssociationPlaneFitting ->
 * accept iff type != NONE && isGoodDepth(depth, 1) (tracking.cc:431-435 / 539-542) -> cam2world with the reference frame's
 * pose (tracking.cc:437-439 / 544-545, camera.cc:149-151).  Outputs (each nullable): type, xyz_depth as in levins_associate;
 * accepted[F]; pw[3 F] world points of accepted features (zeros otherwise; needs pose_ref); undistorted[2 F] pixels. */
typedef struct levins_camera {
    double fx, fy, cx, cy, skew; /* Camera::fx_ ... skew_ */
    double k1, k2, p1, p2, k3;   /* distortion_ in OpenCV order */
} levins_camera;
int levins_associate_pixels(levins_ctx *ctx, const levins_camera *camera, const float *pixels, size_t n_features, int undistort,
                            const levins_pose *pose_ref, int8_t *type, float *xyz_depth, uint8_t *accepted, double *pw,
                            float *undistorted);
/* levins_keyframe with the feature side on the device as above */
int levins_keyframe_pixels(levins_ctx *ctx, double stamp, const levins_ins_sample *ins, size_t n_ins, const levins_camera *camera,
                           const float *pixels, size_t n_features, int undistort, const levins_pose *pose_ref, int8_t *type,
                           float *xyz_depth, uint8_t *accepted, double *pw, float *undistorted, size_t *n_projected_out,
                           size_t *m_out);

/* PointCloudCommon::pointCloudProjection (pointcloud.cc:50-63): dst[k] = f32(R * f64(src[k]) + t). */
int levins_point_cloud_projection(levins_ctx *ctx, const levins_pose *pose, const levins_point_xyzi *src,
                                  levins_point_xyzi *dst, size_t n);

/* ------------------------------------------------------------------------------------------------
 * batched plane geometry of lidar::PointCloudCommon (no call site in the reference, kept for the surface)
 * and the consumer of the depths in the optimiser (rank 4).  Host arrays in and out; one thread per group / factor.
 * ------------------------------------------------------------------------------------------------ */
/* PointCloudCommon::planeEstimation (pointcloud.cc:73-105) on n_groups groups of PLANE_ESTIMATION_POINT_SZIE = 5 points:
 * valid[n], unit_normal[3 n], norm_inverse[n], distance[5 n] (filled up to and including the first failing point, like the
 * reference's `distance` vector; n_distance[n] says how many), each nullable. */
int levins_plane_estimation_batch(levins_ctx *ctx, const levins_point_xyzi *points, size_t n_groups, double threshold,
                                  uint8_t *valid, double *unit_normal, double *norm_inverse, double *distance, int32_t *n_distance);
/* PointCloudCommon::pointToPlaneStd (pointcloud.cc:111-190): n_groups x (group_size neighbours, one target) -> sigma[n]. */
int levins_point_to_plane_std_batch(levins_ctx *ctx, const levins_point_xyzi *nearest, size_t group_size,
                                    const levins_point_xyzi *targets, size_t n_groups, double point_std, double *std_out);
/* DepthFactor (factors/depth_factor.h:38-42) as Optimizer::addReprojectionFactors adds it for depth-associated map points
 * (optimizer.cc:659-665: sigma = Tracking::ASSOCIATE_DEPTH_STD, ceres::HuberLoss(1.0)): residual (depth - 1 / invdepth) / sigma,
 * Jacobian d residual / d invdepth, robustified cost 0.5 rho(residual^2).  huber_a <= 0: no loss function.  Each output nullable. */
int levins_depth_factor_batch(levins_ctx *ctx, const double *depth, const double *invdepth, size_t n, double sigma, double huber_a,
                              double *residual, double *jacobian, double *cost);
/* Optimizer::visualOutlierCullingByChi2 (optimizer.cc:752-783) on n residual blocks of `dim` values: cost = r^T r (nullable),
 * outlier = cost > threshold (5.991 in the reference). */
int levins_chi2_culling_batch(levins_ctx *ctx, const double *residuals, size_t n, size_t dim, double threshold, double *cost,
                              uint8_t *outlier);

/* ------------------------------------------------------------------------------------------------
 * batch API -- many independent keyframes resident in HBM (BASELINE config 5 and the bench workload)
 * ------------------------------------------------------------------------------------------------ */
typedef struct levins_keyframe_desc {
    uint64_t pt_begin, pt_count;   /* slice of the point pool holding this keyframe's buffer (time-sorted) */
    uint64_t ins_begin, ins_count; /* slice of the INS pool */
    uint64_t q_begin, q_count;     /* slice of the query pool (in queries, 3 floats each) */
    double stamp;
} levins_keyframe_desc;

/* Copy the pools and descriptors to the device (host -> device).  Replaces any previous batch.  Asynchronous: the caller's
 * arrays must stay untouched until the next call that synchronises the context (levins_batch_download, levins_synchronize). */
int levins_batch_upload(levins_ctx *ctx, const levins_point_xyzit *points, size_t n_points,
                        const levins_ins_sample *ins, size_t n_ins, const float *query_xyz, size_t n_queries,
                        const levins_keyframe_desc *descs, size_t n_keyframes);
/* The same for pools that already live in this GPU's memory (produced there by another stage, or by a generator): nothing is
 * copied, the three pools stay the caller's and must outlive the batch; `descs` is a host array as above. */
int levins_batch_attach_device(levins_ctx *ctx, const levins_point_xyzit *d_points, size_t n_points, const levins_ins_sample *d_ins,
                               size_t n_ins, const float *d_query_xyz, size_t n_queries, const levins_keyframe_desc *descs,
                               size_t n_keyframes);
/* The same with the points as packed 16-byte records: x, y, z and the point's time as the LiDAR drivers deliver it, a u32
 * count of nanoseconds after its scan's stamp (Ouster `t`, Livox `offset_time`; ROS/lidar/lidar_converter.cc:162, :245:
 * time = stamp + t * 1e-9).  These are the fields of a driver record this path reads (intensity is never used), so the upload
 * moves half the bytes of PointXYZIT records; the records are expanded to PointXYZIT on the device.  `scans` tile the pool
 * in order (scan i covers points [first, first + count)); `stamp` is the scan's header stamp already in the estimator's time
 * base (GPS seconds of week).  Keyframe descriptors index the pool exactly as in levins_batch_upload. */
typedef struct levins_packed_point {
    float x, y, z;
    uint32_t dt_ns;
} levins_packed_point;
typedef struct levins_scan_desc {
    uint64_t first, count;
    double stamp;
} levins_scan_desc;
int levins_batch_upload_packed(levins_ctx *ctx, const levins_packed_point *points, size_t n_points,
                               const levins_scan_desc *scans, size_t n_scans, const levins_ins_sample *ins, size_t n_ins,
                               const float *query_xyz, size_t n_queries, const levins_keyframe_desc *descs, size_t n_keyframes);
/* Enqueue the whole batch on the context's stream: per keyframe window search + INS table (one launch for the
 * batch), then per wave of keyframes deskew+bin, cell finalisation and association.  Asynchronous. */
int levins_batch_run(levins_ctx *ctx);
/* Copy results to the host (device -> host, synchronises).  All outputs nullable.  type / xyz_depth / flags are
 * laid out by descriptor q_begin (i.e. parallel to the query pool); n_projected[K], m[K] per keyframe. */
int levins_batch_download(levins_ctx *ctx, int8_t *type, float *xyz_depth, int32_t *flags, uint32_t *n_projected,
                          uint32_t *m);
/* Device addresses of the result arrays of the resident batch (type: int8[n_queries], xyz_depth: float[4*n_queries],
 * flags: int32[n_queries]), valid until the next levins_batch_upload.  For exchange steps that stay on the device
 * (e.g. an NCCL all-gather of the depths issued on the same stream).  xyz_depth and type are one contiguous block:
 * type == (int8_t *) xyz_depth + 16 * nq16 with nq16 = n_queries rounded up to a multiple of 16, so 17 * nq16 bytes from
 * xyz_depth carry both (and blocks of several ranks can be laid back to back with every float4 aligned). */
int levins_batch_result_pointers(levins_ctx *ctx, void **type, void **xyz_depth, void **flags);
/* Bytes of one rank's result slot for n_queries queries: 17 * (n_queries rounded up to a multiple of 16), always a multiple
 * of 16 (the stride of the slots in a gathered block). */
size_t levins_result_slot_bytes(size_t n_queries);
/* Waves: how many keyframes share one set of launches (their z-buffers stay L2-resident). 0 = default. */
int levins_batch_set_wave(levins_ctx *ctx, int keyframes_per_wave);
/* Per-kernel device time of the last levins_batch_run, measured with CUDA events on the launching stream when
 * profiling is on.  ms[0] prepare, ms[1] deskew+bin, ms[2] plane fit, ms[3] k-NN; launches[i] = number
 * of launches of that kernel. */
int levins_set_profiling(levins_ctx *ctx, int enabled);
int levins_batch_kernel_ms(levins_ctx *ctx, float ms[4], uint32_t launches[4]);
/* number of kernel launches issued by this context since creation */
uint64_t levins_launch_count(const levins_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * exchange step of a sharded batch.  Keyframes are independent, so ranks (one process per GPU, one context
 * each) share nothing until the end, when every rank wants the per-feature results of all ranks.  A rank's results are
 * one contiguous block of 17 bytes per query (xyz_depth, then type; levins_batch_result_pointers; the query count is rounded
 * up to a multiple of 16); the gathered block holds one such slot per rank, in rank order.  Every rank must hold the same
 * number of queries (pad the last shard).
 * The reference is a single process and has no counterpart.  Two interchangeable ways:
 *
 * (1) NCCL: levins_gather enqueues one ncclAllGather on the context's stream.  `comm` is an ncclComm_t -- the caller's own,
 *     or one made by levins_comm_init from an id that rank 0 got from levins_comm_unique_id and handed to the other ranks by
 *     whatever means the host program has.  libnccl.so.2 is resolved at run time; there is no link-time dependency.
 *
 * (2) peer mode: no collective at all.  After levins_peer_export / levins_peer_attach the plane-fit kernel of
 *     levins_batch_run stores each result straight into slot `rank` of EVERY rank's gathered block (peer memory mapped
 *     through CUDA IPC; the stores travel over NVLink / NVSwitch) and publishes a per-rank step counter; levins_peer_wait
 *     holds this context's stream until all ranks have published the current step.  Handles are 64-byte
 *     cudaIpcMemHandle_t images; the host program carries them between the ranks (all ranks call export, exchange the
 *     handles, call attach, then synchronise among themselves once before the first levins_batch_run).  Every rank must
 *     run the same sequence of levins_batch_run calls.  The gathered block is double-buffered by step parity: the results
 *     of step s stay readable until the rank has called levins_peer_wait for step s + 1.
 * ------------------------------------------------------------------------------------------------ */
int levins_comm_unique_id(void *out128);
int levins_comm_init(levins_ctx *ctx, const void *id128, int n_ranks, int rank, void **comm_out);
int levins_comm_destroy(void *comm);
int levins_gather(levins_ctx *ctx, void *nccl_comm, int n_ranks);

int levins_peer_export(levins_ctx *ctx, int n_ranks, size_t n_queries_cap, void *handle_out64);
int levins_peer_attach(levins_ctx *ctx, int rank, int n_ranks, const void *all_handles /* n_ranks x 64 bytes */);
int levins_peer_wait(levins_ctx *ctx);
int levins_peer_detach(levins_ctx *ctx);

/* The gathered block of the latest step on the device (slot r at device_block + r * slot_bytes: xyz_depth
 * float[4 * queries_per_slot], then type int8[queries_per_slot]) and its copy to the host (synchronises; type[n_ranks *
 * queries_per_slot], xyz_depth[4 * n_ranks * queries_per_slot], either nullable). */
int levins_gathered_pointer(levins_ctx *ctx, void **device_block, size_t *slot_bytes, size_t *queries_per_slot);
int levins_gathered_download(levins_ctx *ctx, int8_t *type, float *xyz_depth);

#ifdef __cplusplus
}
#endif
#endif /* LEVINS_ABI_H */
