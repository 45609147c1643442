// ROS/lidar/lidar_converter.h of the reference (:36-62) without the ROS message types: the same class and method names over
// plain record arrays.  Two ways to use it:
//   * the *Convertion methods are the reference's loops on the CPU (kept so that code that needs the PointXYZIT cloud on
//     the host still compiles, and as a second restatement the tests compare with the oracle);
//   * the *Append methods hand the message's own records to the device (levins_append_raw): conversion, range gate and
//     VINS::addNewPointCloud in one call, no PointXYZIT cloud on the host (ROS/fusion.cc:357-376 becomes one line).
#ifndef LEVINS_HOST_LIDAR_CONVERTER_H
#define LEVINS_HOST_LIDAR_CONVERTER_H

#include <cfloat>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "common/device.h"
#include "common/gpstime.h"
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
        nearest_distance_         = nearest_distance;
        farthest_distance_        = farthest_distance;
        nearest_square_distance_  = nearest_distance * nearest_distance;
        farthest_square_distance_ = farthest_distance * farthest_distance;
    }

    // ---- the reference's loops (lidar_converter.cc:35-81, 235-251) ----
    size_t livoxPointCloudConvertion(const livox_ros_driver::CustomMsg &msg, PointCloudCustomPtr &pointcloud_raw,
                                     PointCloudCustomPtr &pointcloud_ds, double &start, double &end, bool to_gps_time) {
        pointcloud_raw->clear();
        pointcloud_raw->reserve(msg.points.size());
        pointcloud_ds->clear();
        pointcloud_ds->reserve(msg.points.size());
        uint64_t timebase = msg.timebase_ns;
        int num_points = 0;
        int filter_num = point_filter_num_;
        start = DBL_MAX;
        end   = 0;
        for (size_t k = 0; k < msg.points.size(); k++) {
            const auto &raw = msg.points[k];
            if ((raw.line < scan_line_) && ((raw.tag & 0x30) == 0x00 || (raw.tag & 0x30) == 0x10)) {
                PointTypeCustom point = livoxPointConvertion(raw, timebase, to_gps_time);
                if (point.time < start) start = point.time;
                if (point.time > end) end = point.time;
                double square_dist = squaredNorm(point);
                if ((square_dist > nearest_square_distance_) && (square_dist < farthest_square_distance_)) {
                    pointcloud_raw->push_back(point);
                    num_points++;
                    if (num_points % filter_num == 0) pointcloud_ds->push_back(point);
                }
            }
        }
        return pointcloud_ds->size();
    }

    static PointTypeCustom livoxPointConvertion(const livox_ros_driver::CustomPoint &raw, uint64_t timebase, bool to_gps_time) {
        PointTypeCustom point;
        point.data3     = 0; // (uninitialised in the reference; zero here like the device path)
        point.x         = raw.x;
        point.y         = raw.y;
        point.z         = raw.z;
        point.intensity = raw.reflectivity;
        point.time      = static_cast<double>(timebase + raw.offset_time) * 1.0e-9;
        if (to_gps_time) {
            int week;
            double sow;
            GpsTime::unix2gps(point.time, week, sow);
            point.time = sow;
        }
        return point;
    }

    // lidar_converter.cc:83-130 / 132-179 / 181-233: one template for the three PointCloud2 drivers
    size_t velodynePointCloudConvertion(const std::vector<VelodynePoint> &raw, double stamp, PointCloudCustomPtr &pointcloud,
                                        double &start, double &end, bool to_gps_time) {
        return cloud2Convertion(raw, stamp, pointcloud, start, end, to_gps_time,
                                [](const VelodynePoint &p, double s, bool) { return s + p.time; }, true);
    }
    size_t ousterPointCloudConvertion(const std::vector<OusterPoint> &raw, double stamp, PointCloudCustomPtr &pointcloud,
                                      double &start, double &end, bool to_gps_time) {
        return cloud2Convertion(raw, stamp, pointcloud, start, end, to_gps_time,
                                [](const OusterPoint &p, double s, bool) { return s + p.t * 1.0e-9; }, true);
    }
    size_t hesaiPointCloudConvertion(const std::vector<HesaiPoint> &raw, double stamp, PointCloudCustomPtr &pointcloud, double &start,
                                     double &end, bool to_gps_time) {
        return cloud2Convertion(raw, stamp, pointcloud, start, end, to_gps_time,
                                [](const HesaiPoint &p, double, bool gps) {
                                    if (!gps) return p.timestamp;
                                    int week;
                                    double sow;
                                    GpsTime::unix2gps(p.timestamp, week, sow);
                                    return sow;
                                },
                                false);
    }

    // ---- the same conversions on the device, straight into the ring buffer of the depth-enhancement path ----
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
    static double squaredNorm(const PointTypeCustom &p) { return static_cast<double>(p.x * p.x + (p.y * p.y + p.z * p.z)); } // Vector3f

    template <typename Raw, typename TimeFn>
    size_t cloud2Convertion(const std::vector<Raw> &pointcloud_raw, double stamp, PointCloudCustomPtr &pointcloud, double &start,
                            double &end, bool to_gps_time, TimeFn time_of, bool stamp_relative) {
        pointcloud->clear();
        if (stamp_relative && to_gps_time) {
            int week;
            double sow;
            GpsTime::unix2gps(stamp, week, sow);
            stamp = sow;
        }
        std::vector<int> point_counts(scan_line_, 0);
        PointTypeCustom point;
        point.data3 = 0;
        start       = DBL_MAX;
        end         = 0;
        for (size_t k = 0; k < pointcloud_raw.size(); k++) {
            auto &raw = pointcloud_raw[k];
            if (point_counts[raw.ring] % point_filter_num_ == 0) {
                point.x         = raw.x;
                point.y         = raw.y;
                point.z         = raw.z;
                point.intensity = raw.intensity;
                point.time      = time_of(raw, stamp, to_gps_time);
                if (point.time < start) start = point.time;
                if (point.time > end) end = point.time;
                double square_dist = squaredNorm(point);
                if ((square_dist > nearest_square_distance_) && (square_dist < farthest_square_distance_)) pointcloud->push_back(point);
            }
            point_counts[raw.ring]++;
        }
        return pointcloud->size();
    }

    size_t append(const levins_raw_format &fmt, const void *records, size_t n, uint64_t timebase, double stamp, double &start,
                  double &end, bool to_gps_time) {
        size_t appended = 0;
        levins_host::check(levins_append_raw(levins_host::context(), &fmt, records, n, timebase, stamp, to_gps_time ? 1 : 0, scan_line_,
                                             nearest_distance_, farthest_distance_, &appended, &start, &end),
                           "levins_append_raw");
        return appended;
    }

    int scan_line_{6};
    int point_filter_num_{1};
    double nearest_distance_{1.0}, farthest_distance_{200.0};
    double nearest_square_distance_{1.0};
    double farthest_square_distance_{200.0 * 200.0};
};

#endif
