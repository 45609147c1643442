// lidar::PointCloudCommon of the reference (le_vins/lidar/pointcloud.h:29-52): same statics.  The single-point
// projections are trivial host inlines (they exist for API compatibility; the per-point loop that called them at
// le_vins.cc:461 is k_deskew_bin); the whole-cloud projection (pointcloud.cc:50-63, tbb::parallel_for) runs on the GPU.
#ifndef LEVINS_HOST_POINTCLOUD_H
#define LEVINS_HOST_POINTCLOUD_H

#include <cstring>

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

public:
    static constexpr int PLANE_ESTIMATION_POINT_SZIE = 5;
};

} // namespace lidar
#endif
