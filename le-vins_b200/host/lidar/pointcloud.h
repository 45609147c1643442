// lidar::PointCloudCommon of the reference (le_vins/lidar/pointcloud.h:29-52): same statics.  The single-point
// projections are trivial host inlines (they exist for API compatibility; the per-point loop that called them at
// le_vins.cc:461 is k_deskew_bin); the whole-cloud projection (pointcloud.cc:50-63, tbb::parallel_for) runs on the GPU.
#ifndef LEVINS_HOST_POINTCLOUD_H
#define LEVINS_HOST_POINTCLOUD_H

#include <cstring>

#include <vector>

#include "common/device.h"
#include "common/types.h"
#include "lidar/lidar.h"

namespace lidar {

class PointCloudCommon {
public:
    static PointType pointProjection(const Pose &pose, const PointType &point) { // pointcloud.cc:65-71
        PointType projected = point;
        const Vector3d r = pose.R * Vector3d(point.x, point.y, point.z) + pose.t;
        projected.x = static_cast<float>(r[0]);
        projected.y = static_cast<float>(r[1]);
        projected.z = static_cast<float>(r[2]);
        return projected;
    }

    static PointType relativeProjection(const Pose &pose0, const Pose &pose1, const PointType &point_l1) { // :28-34
        Pose rel;
        rel.R = pose0.R.transpose() * pose1.R;
        rel.t = pose0.R.transpose() * (pose1.t - pose0.t);
        return pointProjection(rel, point_l1);
    }

    static PointType relativeProjectionToCamera(const Pose &pose_c_l, const Pose &pose_cam0, const Pose &pose_cam1,
                                                const PointType &point_lidar) { // :36-48
        const PointType point_cam1 = pointProjection(pose_c_l, point_lidar);
        Pose rel;
        rel.R = pose_cam0.R.transpose() * pose_cam1.R;
        rel.t = pose_cam0.R.transpose() * (pose_cam1.t - pose_cam0.t);
        return pointProjection(rel, point_cam1);
    }

    // GPU: levins_point_cloud_projection (k_project)
    static void pointCloudProjection(const Pose &pose, PointCloudPtr src, PointCloudPtr dst) { // :50-63
        dst->resize(src->size());
        levins_pose p;
        std::memcpy(p.R, pose.R.data(), sizeof(p.R));
        std::memcpy(p.t, pose.t.data(), sizeof(p.t));
        levins_host::check(levins_point_cloud_projection(levins_host::context(), &p,
                                                         reinterpret_cast<const levins_point_xyzi *>(src->points.data()),
                                                         reinterpret_cast<levins_point_xyzi *>(dst->points.data()), src->size()),
                           "pointCloudProjection");
    }

    static double pointSquareDistance(const PointType &point0, const PointType &point1) { // :107-109
        const float dx = point0.x - point1.x, dy = point0.y - point1.y, dz = point0.z - point1.z;
        return dx * dx + (dy * dy + dz * dz);
    }

    // pointcloud.cc:73-105 (no call site in the reference; kept for the drop-in surface).  GPU: k_plane_estimation, one thread
    // per group of 5 points; the single-group form is the batch of one.
    static bool planeEstimation(const PointVector &points, double threshold, Vector3d &unit_normal_vector, double &norm_inverse,
                                std::vector<double> &distance) {
        if (points.size() < (size_t) PLANE_ESTIMATION_POINT_SZIE) return false;
        uint8_t valid = 0;
        double un[3], dist[5];
        int32_t nd = 0;
        levins_host::check(levins_plane_estimation_batch(levins_host::context(), reinterpret_cast<const levins_point_xyzi *>(points.data()), 1,
                                                         threshold, &valid, un, &norm_inverse, dist, &nd),
                           "planeEstimation");
        unit_normal_vector = Vector3d(un[0], un[1], un[2]);
        distance.assign(dist, dist + nd);
        return valid != 0;
    }
    // many groups at once: points holds 5 consecutive points per group
    static void planeEstimationBatch(const PointVector &points, double threshold, std::vector<uint8_t> &valid,
                                     std::vector<double> &unit_normals, std::vector<double> &norm_inverses) {
        const size_t n = points.size() / PLANE_ESTIMATION_POINT_SZIE;
        valid.resize(n);
        unit_normals.resize(3 * n);
        norm_inverses.resize(n);
        levins_host::check(levins_plane_estimation_batch(levins_host::context(), reinterpret_cast<const levins_point_xyzi *>(points.data()), n,
                                                         threshold, valid.data(), unit_normals.data(), norm_inverses.data(), nullptr, nullptr),
                           "planeEstimationBatch");
    }

    // pointcloud.cc:111-190 (no call site in the reference): point-to-plane sigma by first-order covariance propagation
    // through the eigen-decomposition of the neighbourhood covariance.  GPU: k_point_to_plane_std.
    static double pointToPlaneStd(const PointVector &nearest, const PointType &target, double point_std) {
        double out = 0;
        levins_host::check(levins_point_to_plane_std_batch(levins_host::context(), reinterpret_cast<const levins_point_xyzi *>(nearest.data()),
                                                           nearest.size(), reinterpret_cast<const levins_point_xyzi *>(&target), 1, point_std,
                                                           &out),
                           "pointToPlaneStd");
        return out;
    }
    static void pointToPlaneStdBatch(const PointVector &nearest, size_t group_size, const PointVector &targets, double point_std,
                                     std::vector<double> &stds) {
        stds.resize(targets.size());
        levins_host::check(levins_point_to_plane_std_batch(levins_host::context(), reinterpret_cast<const levins_point_xyzi *>(nearest.data()),
                                                           group_size, reinterpret_cast<const levins_point_xyzi *>(targets.data()),
                                                           targets.size(), point_std, stds.data()),
                           "pointToPlaneStdBatch");
    }

public:
    static constexpr int PLANE_ESTIMATION_POINT_SZIE = 5;
};

} // namespace lidar
#endif
