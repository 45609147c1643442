// The LiDAR-enhancement slice of class VINS (le_vins/le_vins/le_vins.h:65,109,130,166,183-184):
//   VINS::addNewPointCloud                      le_vins.cc:149-163
//   VINS::pointcloudStackForVisualTracking      le_vins.cc:392-470
// Same method names, argument meaning and bool returns.  What the real class reads from its members (frame stamp,
// ins_window_, pose_b_c_, pose_c_l_, td_b_l_) is passed explicitly, because the estimator that owns those members is
// out of scope; INTEGRATION.md shows the two method bodies a maintainer replaces.  All arithmetic runs in
// liblevins_b200.so (CUDA); nothing here computes a point on the CPU.
#ifndef LEVINS_HOST_VINS_LIDAR_H
#define LEVINS_HOST_VINS_LIDAR_H

#include <atomic>
#include <cstring>
#include <mutex>
#include <unistd.h>
#include <vector>

#include "common/device.h"
#include "common/misc.h"
#include "lidar/pointcloud.h"

namespace levins_host {

// which host PointCloud mirrors the deskewed cloud that is still resident on the device (set by the stacker, read by
// Tracking::depthAssociationProjection to skip the re-upload)
struct ResidentCloud {
    const void *ptr = nullptr;
    size_t size = 0;
};
inline ResidentCloud &resident_cloud() {
    static ResidentCloud r;
    return r;
}

class VinsLidarPath {
public:
    static constexpr double LIDAR_DEPTH_POINT_STACK_INTERVAL = 0.5; // le_vins.h:130
    static constexpr int WAITING_DELAY_IN_US = 100;

    VinsLidarPath(const Pose &pose_b_c, const Pose &pose_c_l, double td_b_l) {
        levins_pose b, c;
        to_abi(pose_b_c, b);
        to_abi(pose_c_l, c);
        check(levins_set_extrinsics(context(), &b, &c, td_b_l), "set_extrinsics");
    }

    // le_vins.cc:149-163 (called from the LiDAR callback thread in the reference; thread-safe here too)
    bool addNewPointCloud(PointCloudCustomPtr pointcloud) {
        if (!is_use_lidar_depth_ || pointcloud->empty()) return true;
        check(levins_append_scan(context(), reinterpret_cast<const levins_point_xyzit *>(pointcloud->points.data()),
                                 pointcloud->size()),
              "addNewPointCloud");
        newest_point_time_.store(pointcloud->points.back().time);
        have_points_.store(true);
        return true;
    }

    // a scan that went to the device as driver records (LidarConverter::*PointCloudAppend): `newest_time` = the converter's `end`
    void notifyAppended(size_t appended, double newest_time) {
        if (!appended) return;
        newest_point_time_.store(newest_time);
        have_points_.store(true);
    }

    bool lidarPointBufferEmpty() const { return !have_points_.load(); }

    // le_vins.cc:392-470.  frame_stamp = frame->stamp(); ins_window = ins_window_.
    bool pointcloudStackForVisualTracking(double frame_stamp, const MISC::InsWindow &ins_window, PointCloud &pointcloud) {
        while (!is_finished_) { // le_vins.cc:398-407: wait until the newest LiDAR point is later than the frame
            if (newest_point_time_.load() > frame_stamp) break;
            usleep(WAITING_DELAY_IN_US);
        }
        to_abi(ins_window, ins_);
        size_t cap = 0;
        check(levins_buffer_info(context(), &cap, nullptr, nullptr), "buffer_info");
        pointcloud.points.resize(cap);
        size_t n = 0;
        const int st = levins_stack_deskew(context(), frame_stamp, ins_.data(), ins_.size(),
                                           reinterpret_cast<levins_point_xyzi *>(pointcloud.points.data()), cap, &n);
        check(st, "pointcloudStackForVisualTracking");
        pointcloud.points.resize(n);
        resident_cloud().ptr = &pointcloud;
        resident_cloud().size = n;
        return st == LEVINS_OK && !pointcloud.empty(); // le_vins.cc:469
    }

    void setFinished() { is_finished_ = true; }

private:
    bool is_use_lidar_depth_{true};
    std::atomic<bool> is_finished_{false};
    std::atomic<bool> have_points_{false};
    std::atomic<double> newest_point_time_{-1e300};
    std::vector<levins_ins_sample> ins_;
};

} // namespace levins_host
#endif
