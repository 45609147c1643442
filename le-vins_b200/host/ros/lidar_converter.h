// ROS/lidar/lidar_converter.h of the reference (:36-62) without the ROS message types: the same class over plain record
// arrays, with the conversion moved to the device.  Each *PointCloudAppend method hands the message's own records to
// levins_append_raw: LidarConverter::*PointCloudConvertion (ROS/lidar/lidar_converter.cc:35-256: tag / line filter, time
// conversion, range gate, start / end) and VINS::addNewPointCloud (le_vins.cc:149-163) in one call, no PointXYZIT cloud on
// the host (ROS/fusion.cc:357-376 becomes one line).  There is no CPU conversion in this header.
#ifndef LEVINS_HOST_LIDAR_CONVERTER_H
#define LEVINS_HOST_LIDAR_CONVERTER_H

#include <cfloat>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "common/device.h"
#include "lidar/lidar.h"

namespace livox_ros_driver {
// livox_ros_driver/CustomPoint.msg as the generated C++ struct lays it out (20 bytes, 1 byte of tail padding)
struct CustomPoint {
    uint32_t offset_time;
    float x, y, z;
    uint8_t reflectivity, tag, line;
};
static_assert(sizeof(CustomPoint) == 20, "CustomPoint");
struct CustomMsg {
    uint64_t timebase_ns = 0; // msg->header.stamp.toNSec()
    std::vector<CustomPoint> points;
};
} // namespace livox_ros_driver

// the pcl point structs PointCloud2 messages are unpacked into (le_vins/lidar/lidar.h:36-76)
struct alignas(16) VelodynePoint {
    float x, y, z, data3;
    float intensity;
    float time;
    uint16_t ring;
};
struct alignas(16) OusterPoint {
    float x, y, z, data3;
    float intensity;
    uint32_t t;
    uint16_t reflectivity, ambient;
    uint32_t range;
    uint8_t ring;
};
struct alignas(16) HesaiPoint {
    float x, y, z, data3;
    float intensity;
    double timestamp;
    uint16_t ring;
};
static_assert(sizeof(VelodynePoint) == 32 && sizeof(OusterPoint) == 48 && sizeof(HesaiPoint) == 48, "pcl point structs");

class LidarConverter {
public:
    LidarConverter(double frame_rate, int scan_line, double nearest_distance, double farthest_distance) : scan_line_(scan_line) {
        (void) frame_rate;
        nearest_distance_  = nearest_distance;
        farthest_distance_ = farthest_distance;
    }

    // ---- conversion + append on the device, straight into the ring buffer of the depth-enhancement path ----
    size_t livoxPointCloudAppend(const livox_ros_driver::CustomMsg &msg, double &start, double &end, bool to_gps_time) {
        static const levins_raw_format fmt = {1, sizeof(livox_ros_driver::CustomPoint), 4, 8, 12, 16, 0, 18, 17, 0, 1, 0};
        return append(fmt, msg.points.data(), msg.points.size(), msg.timebase_ns, 0.0, start, end, to_gps_time);
    }
    size_t velodynePointCloudAppend(const std::vector<VelodynePoint> &raw, double stamp, double &start, double &end, bool to_gps_time) {
        static const levins_raw_format fmt = {2, sizeof(VelodynePoint), 0, 4, 8, 16, 20, 24, 0, 1, 2, 0};
        return append(fmt, raw.data(), raw.size(), 0, stamp, start, end, to_gps_time);
    }
    size_t ousterPointCloudAppend(const std::vector<OusterPoint> &raw, double stamp, double &start, double &end, bool to_gps_time) {
        static const levins_raw_format fmt = {2, sizeof(OusterPoint), 0, 4, 8, 16, 20, 32, 0, 2, 1, 0};
        return append(fmt, raw.data(), raw.size(), 0, stamp, start, end, to_gps_time);
    }
    size_t hesaiPointCloudAppend(const std::vector<HesaiPoint> &raw, double stamp, double &start, double &end, bool to_gps_time) {
        static const levins_raw_format fmt = {2, sizeof(HesaiPoint), 0, 4, 8, 16, 24, 32, 0, 3, 2, 0};
        return append(fmt, raw.data(), raw.size(), 0, stamp, start, end, to_gps_time);
    }

private:
    size_t append(const levins_raw_format &fmt, const void *records, size_t n, uint64_t timebase, double stamp, double &start,
                  double &end, bool to_gps_time) {
        size_t appended = 0;
        levins_host::check(levins_append_raw(levins_host::context(), &fmt, records, n, timebase, stamp, to_gps_time ? 1 : 0, scan_line_,
                                             nearest_distance_, farthest_distance_, &appended, &start, &end),
                           "levins_append_raw");
        return appended;
    }

    int scan_line_{6};
    double nearest_distance_{1.0}, farthest_distance_{200.0};
};

#endif
