// The depth-association slice of visual::Tracking (le_vins/visual/tracking.h:92-98, 119-135, 159-160):
//   bool depthAssociationProjection(PointCloud &)            tracking.cc:298-368
//   MapPointType depthAssociationPlaneFitting(PointType &)    tracking.cc:564-634
//   int depthAssociationReference() / depthAssociationMapPoint()   tracking.cc:370-562 (drivers)
// Same names and return conventions.  The drivers work on flat arrays: creation of VisualMapPoint / VisualFeature
// objects (tracking.cc:437-465, 544-549) stays in the estimator, which owns that object graph.
#ifndef LEVINS_HOST_TRACKING_DEPTH_H
#define LEVINS_HOST_TRACKING_DEPTH_H

#include <cstdint>
#include <vector>

#include "common/device.h"
#include "common/misc.h"
#include "le_vins/vins_lidar.h"
#include "lidar/lidar.h"
#include "visual/camera.h"
#include "visual/visual_mappoint.h"

namespace visual {

struct DepthAssociated { // one accepted feature: what createMapPoint / setReferenceFrame need (tracking.cc:437-440, 544-549)
    size_t index;        // index into the driver's input list
    Vector3d pw;         // world point, camera_->cam2world(pc, frame_ref_->pose())
    double depth;        // lidar depth (= z in the reference frame's camera)
    MapPointType type;
};

class Tracking {
public:
    static constexpr double ASSOCIATE_DEPTH_STD = 0.1; // tracking.h:98

    // tracking.cc:298-368.  Returns false when no cell is filled (tracking.cc:356-358).
    bool depthAssociationProjection(PointCloud &pointcloud) {
        levins_ctx *ctx = levins_host::context();
        const levins_host::ResidentCloud &res = levins_host::resident_cloud();
        const bool resident = res.ptr == &pointcloud && res.size == pointcloud.size() && res.size > 0;
        size_t m = 0;
        int st = resident ? levins_build_range_image(ctx, nullptr, 0, nullptr, 0, nullptr, &m) : LEVINS_ERR_STATE;
        if (st == LEVINS_ERR_STATE) // not (or no longer) resident: bin the caller's cloud
            st = levins_build_range_image(ctx, reinterpret_cast<const levins_point_xyzi *>(pointcloud.points.data()),
                                          pointcloud.size(), nullptr, 0, nullptr, &m);
        levins_host::check(st, "depthAssociationProjection");
        unit_sphere_size_ = m;
        return st == LEVINS_OK && m > 0;
    }

    size_t unitSphereSize() const { return unit_sphere_size_; }

    // tracking.cc:564-634 for one query: the point is modified in place like in the reference
    MapPointType depthAssociationPlaneFitting(PointType &point) {
        std::vector<PointType> one(1, point);
        std::vector<MapPointType> type;
        depthAssociationPlaneFitting(one, std::vector<uint8_t>(), type);
        point = one[0];
        return type[0];
    }

    // The per-feature loops of tracking.cc:399-413 / 512-525 as one batched call.  skip[k] != 0 reproduces the
    // `continue` branches (tracking.cc:401-404, 513-516): such features get MAPPOINT_NONE and are not evaluated.
    // Returns num_succeed.
    int depthAssociationPlaneFitting(std::vector<PointType> &unit_sphere_features, const std::vector<uint8_t> &skip,
                                     std::vector<MapPointType> &mappoint_types) {
        const size_t f = unit_sphere_features.size();
        mappoint_types.assign(f, MAPPOINT_NONE);
        if (f == 0 || unit_sphere_size_ == 0) return 0;
        q_.clear();
        map_.clear();
        for (size_t k = 0; k < f; k++) {
            if (!skip.empty() && skip[k]) continue;
            const PointType &p = unit_sphere_features[k];
            q_.insert(q_.end(), {p.x, p.y, p.z});
            map_.push_back(k);
        }
        const size_t n = map_.size();
        type_.resize(n);
        xyzd_.resize(4 * n);
        levins_host::check(levins_associate(levins_host::context(), q_.data(), n, type_.data(), xyzd_.data(), nullptr, nullptr, nullptr),
                           "depthAssociationPlaneFitting");
        int num_succeed = 0;
        for (size_t i = 0; i < n; i++) {
            PointType &p = unit_sphere_features[map_[i]];
            mappoint_types[map_[i]] = static_cast<MapPointType>(type_[i]);
            if (type_[i] != MAPPOINT_NONE) {
                p.x = xyzd_[4 * i];
                p.y = xyzd_[4 * i + 1];
                p.z = xyzd_[4 * i + 2];
                p.intensity = xyzd_[4 * i + 3]; // tracking.cc:629
                num_succeed++;
            }
        }
        return num_succeed;
    }

    // Shared body of depthAssociationReference (tracking.cc:381-465) and depthAssociationMapPoint (489-556), one device call:
    // tracked reference-frame pixels -> [undistort] -> unit rays (f32) -> plane fitting -> gate (type != NONE &&
    // isGoodDepth) -> world point.  `skip` marks features whose reference frame is not frame_ref_ (tracking.cc:401) resp. map
    // points that are already depth associated (tracking.cc:513): they are not evaluated.  `undistort`: the pixels are the raw
    // tracked ones (pts2d_ref_), Camera::undistortPoints runs on the device too; false when the caller undistorted already.
    // Returns num_associated; `associated` lists the accepted ones.
    int depthAssociation(const Camera &camera, const std::vector<Point2f> &pts2d, const std::vector<uint8_t> &skip,
                         const Pose &frame_ref_pose, std::vector<DepthAssociated> &associated, bool undistort = false) {
        associated.clear();
        if (pts2d.empty() || unit_sphere_size_ == 0) return 0; // tracking.cc:372-374 / 483-485, 165
        q_.clear();
        map_.clear();
        for (size_t k = 0; k < pts2d.size(); k++) {
            if (!skip.empty() && skip[k]) continue;
            q_.insert(q_.end(), {pts2d[k].x, pts2d[k].y});
            map_.push_back(k);
        }
        const size_t n = map_.size();
        if (n == 0) return 0;
        type_.resize(n);
        xyzd_.resize(4 * n);
        accepted_.resize(n);
        pw_.resize(3 * n);
        const levins_camera cam = camera.abi();
        levins_pose pose;
        levins_host::to_abi(frame_ref_pose, pose);
        levins_host::check(levins_associate_pixels(levins_host::context(), &cam, q_.data(), n, undistort ? 1 : 0, &pose, type_.data(),
                                                   xyzd_.data(), accepted_.data(), pw_.data(), nullptr),
                           "depthAssociation");
        for (size_t i = 0; i < n; i++)
            if (accepted_[i])
                associated.push_back({map_[i], Vector3d(pw_[3 * i], pw_[3 * i + 1], pw_[3 * i + 2]), (double) xyzd_[4 * i + 3],
                                      static_cast<MapPointType>(type_[i])});
        return static_cast<int>(associated.size());
    }

    // tracking.cc:370-480: the pixels are pts2d_ref_ as tracked (undistort = true runs tracking.cc:381-385 on the device)
    int depthAssociationReference(const Camera &camera, const std::vector<Point2f> &pts2d_ref,
                                  const std::vector<uint8_t> &not_current_ref, const Pose &frame_ref_pose,
                                  std::vector<DepthAssociated> &associated, bool undistort = false) {
        return depthAssociation(camera, pts2d_ref, not_current_ref, frame_ref_pose, associated, undistort);
    }
    // tracking.cc:482-562: the pixels are the map points' reference-frame keypoints (already undistorted, :489-504)
    int depthAssociationMapPoint(const Camera &camera, const std::vector<Point2f> &pts2d_map_undis,
                                 const std::vector<uint8_t> &already_depth_associated, const Pose &frame_ref_pose,
                                 std::vector<DepthAssociated> &associated) {
        return depthAssociation(camera, pts2d_map_undis, already_depth_associated, frame_ref_pose, associated, false);
    }

private:
    size_t unit_sphere_size_ = 0;
    std::vector<float> q_, xyzd_;
    std::vector<int8_t> type_;
    std::vector<uint8_t> accepted_;
    std::vector<double> pw_;
    std::vector<size_t> map_;
};

} // namespace visual
#endif
