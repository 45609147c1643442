// Drives the C++ host mirror exactly like the estimator would (le_vins.cc:367-383, tracking.cc:160-172) on a keyframe
// dumped by tests/test_host_mirror.py, and writes its outputs back for comparison with the oracle.
//   test_host <dir>   reads  points.bin ins.bin pixels.bin meta.bin   writes  cloud.bin assoc.bin summary.txt
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "le_vins/vins_lidar.h"
#include "lidar/pointcloud.h"
#include "ros/lidar_converter.h"
#include "visual/tracking_depth.h"

template <typename T> static std::vector<T> read_bin(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    const size_t bytes = f.tellg();
    f.seekg(0);
    std::vector<T> v(bytes / sizeof(T));
    f.read(reinterpret_cast<char *>(v.data()), bytes);
    return v;
}
template <typename T> static void write_bin(const std::string &path, const std::vector<T> &v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), v.size() * sizeof(T));
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    try {
        // meta: stamp, td_b_l, pose_b_c[12], pose_c_l[12], intr[5], frame_ref_pose[12], scans
        const auto meta = read_bin<double>(dir + "/meta.bin");
        const double stamp = meta[0], td_b_l = meta[1];
        Pose pose_b_c, pose_c_l, frame_pose;
        std::memcpy(pose_b_c.R.data(), &meta[2], 72);
        std::memcpy(pose_b_c.t.data(), &meta[11], 24);
        std::memcpy(pose_c_l.R.data(), &meta[14], 72);
        std::memcpy(pose_c_l.t.data(), &meta[23], 24);
        visual::Camera camera(meta[26], meta[27], meta[28], meta[29], meta[30]);
        std::memcpy(frame_pose.R.data(), &meta[31], 72);
        std::memcpy(frame_pose.t.data(), &meta[40], 24);
        const int n_scans = static_cast<int>(meta[43]);

        const auto points = read_bin<PointXYZIT>(dir + "/points.bin");
        const auto ins_raw = read_bin<double>(dir + "/ins.bin"); // 9 doubles per sample
        const auto pixels = read_bin<float>(dir + "/pixels.bin");

        MISC::InsWindow ins_window;
        for (size_t i = 0; i + 8 < ins_raw.size(); i += 9) {
            IMU imu;
            imu.time = ins_raw[i];
            IntegrationState st;
            st.time = ins_raw[i + 1];
            st.p = Vector3d(ins_raw[i + 2], ins_raw[i + 3], ins_raw[i + 4]);
            st.q = Quaterniond(&ins_raw[i + 5]);
            ins_window.push_back({imu, st});
        }

        levins_host::VinsLidarPath vins(pose_b_c, pose_c_l, td_b_l);
        // the LiDAR callback: one addNewPointCloud per scan (fusion.cc:360 / 376)
        const size_t per = (points.size() + n_scans - 1) / n_scans;
        for (size_t b = 0; b < points.size(); b += per) {
            PointCloudCustomPtr scan(new PointCloudCustom);
            scan->points.assign(points.begin() + b, points.begin() + std::min(points.size(), b + per));
            vins.addNewPointCloud(scan);
        }
        // le_vins.cc:376-381
        static PointCloud pointcloud;
        pointcloud.clear();
        const bool stacked = vins.pointcloudStackForVisualTracking(stamp, ins_window, pointcloud);

        // tracking.cc:160-172
        visual::Tracking tracking;
        std::vector<visual::DepthAssociated> associated;
        int num_associated = 0;
        bool projected = false;
        if (!pointcloud.empty()) {
            projected = tracking.depthAssociationProjection(pointcloud);
            if (projected) {
                std::vector<visual::Point2f> px(pixels.size() / 2);
                for (size_t k = 0; k < px.size(); k++) px[k] = {pixels[2 * k], pixels[2 * k + 1]};
                num_associated = tracking.depthAssociationReference(camera, px, std::vector<uint8_t>(), frame_pose, associated);
            }
        }
        // the drivers' skip branches: features whose reference frame is not frame_ref_ (tracking.cc:401-404) and map points
        // that are already depth associated (tracking.cc:513-516) are not evaluated
        if (projected) {
            std::vector<visual::Point2f> px(pixels.size() / 2);
            for (size_t k = 0; k < px.size(); k++) px[k] = {pixels[2 * k], pixels[2 * k + 1]};
            std::vector<uint8_t> not_current_ref(px.size()), already(px.size());
            for (size_t k = 0; k < px.size(); k++) {
                not_current_ref[k] = (k % 3 == 0);
                already[k] = (k % 4 == 1);
            }
            auto dump = [&](const char *name, const std::vector<visual::DepthAssociated> &list, int count) {
                std::vector<double> o{(double) count};
                for (const auto &a : list) {
                    o.push_back((double) a.index);
                    o.push_back(a.depth);
                    o.push_back((double) a.type);
                }
                write_bin(dir + "/" + name, o);
            };
            std::vector<visual::DepthAssociated> a_ref, a_map;
            const int n_ref = tracking.depthAssociationReference(camera, px, not_current_ref, frame_pose, a_ref);
            dump("assoc_ref_skip.bin", a_ref, n_ref);
            const int n_map = tracking.depthAssociationMapPoint(camera, px, already, frame_pose, a_map);
            dump("assoc_map_skip.bin", a_map, n_map);
        }
        // single-query form on the first pixel (in-place semantics of tracking.cc:626-629)
        PointType single;
        int single_type = -1;
        if (projected && !pixels.empty()) {
            const Vector3d puc = camera.pixel2unitcam({pixels[0], pixels[1]});
            single.x = (float) puc[0];
            single.y = (float) puc[1];
            single.z = (float) puc[2];
            single_type = tracking.depthAssociationPlaneFitting(single);
        }

        write_bin(dir + "/cloud.bin", pointcloud.points);
        std::vector<double> out;
        for (const auto &a : associated) {
            out.push_back((double) a.index);
            out.push_back(a.pw[0]);
            out.push_back(a.pw[1]);
            out.push_back(a.pw[2]);
            out.push_back(a.depth);
            out.push_back((double) a.type);
        }
        write_bin(dir + "/assoc.bin", out);
        std::ofstream s(dir + "/summary.txt");
        s << "stacked " << stacked << "\nn " << pointcloud.size() << "\nprojected " << projected << "\nM " << tracking.unitSphereSize()
          << "\nassociated " << num_associated << "\nsingle_type " << single_type << "\nsingle " << single.x << " " << single.y << " "
          << single.z << " " << single.intensity << "\n";
        std::printf("host mirror: n=%zu M=%zu associated=%d\n", pointcloud.size(), tracking.unitSphereSize(), num_associated);

        // the same scans once more as the driver's own Livox records: LidarConverter::livoxPointCloudAppend converts them on
        // the device (ROS/fusion.cc:357-360 in one call), then the stacker runs again
        {
            std::ifstream probe(dir + "/raw_livox.bin");
            if (probe) {
                const auto recs = read_bin<livox_ros_driver::CustomPoint>(dir + "/raw_livox.bin");
                const auto tb = read_bin<uint64_t>(dir + "/raw_timebase.bin"); // per scan: timebase, record count
                levins_host::check(levins_buffer_clear(levins_host::context()), "buffer_clear");
                LidarConverter converter(10.0, 6, 1.0, 200.0);
                size_t o = 0, total = 0;
                for (size_t sidx = 0; sidx + 1 < tb.size(); sidx += 2) {
                    livox_ros_driver::CustomMsg msg;
                    msg.timebase_ns = tb[sidx];
                    msg.points.assign(recs.begin() + o, recs.begin() + o + tb[sidx + 1]);
                    o += tb[sidx + 1];
                    double start = 0, end = 0;
                    const size_t n = converter.livoxPointCloudAppend(msg, start, end, true);
                    vins.notifyAppended(n, end);
                    total += n;
                }
                PointCloud from_records;
                const bool ok = vins.pointcloudStackForVisualTracking(stamp, ins_window, from_records);
                write_bin(dir + "/cloud_raw.bin", from_records.points);
                std::ofstream s2(dir + "/summary_raw.txt");
                s2 << "stacked " << ok << "\nappended " << total << "\nn " << from_records.size() << "\n";
            }
        }
        // PointCloudCommon::planeEstimation / pointToPlaneStd / pointSquareDistance (GPU-backed) on consecutive groups of 5 points
        {
            std::ifstream probe(dir + "/pts.bin");
            if (probe) {
                const auto pts = read_bin<PointType>(dir + "/pts.bin");
                std::vector<double> plane, stds;
                for (size_t k = 0; k + 5 <= pts.size(); k += 5) {
                    PointVector five(pts.begin() + k, pts.begin() + k + 5);
                    Vector3d n;
                    double ninv = 0;
                    std::vector<double> dist;
                    const bool ok = lidar::PointCloudCommon::planeEstimation(five, 0.1, n, ninv, dist);
                    plane.insert(plane.end(), {ok ? 1.0 : 0.0, n[0], n[1], n[2], ninv, (double) dist.size()});
                    stds.push_back(lidar::PointCloudCommon::pointToPlaneStd(five, pts[(k + 7) % pts.size()], 0.03));
                    stds.push_back(lidar::PointCloudCommon::pointSquareDistance(pts[k], pts[k + 1]));
                }
                write_bin(dir + "/plane.bin", plane);
                write_bin(dir + "/std.bin", stds);
                std::vector<uint8_t> valid;
                std::vector<double> normals, ninvs;
                lidar::PointCloudCommon::planeEstimationBatch(PointVector(pts.begin(), pts.end()), 0.1, valid, normals, ninvs);
                std::vector<double> vb(valid.begin(), valid.end());
                write_bin(dir + "/plane_batch.bin", vb);
            }
        }
        // the three PointCloud2 drivers through LidarConverter::*PointCloudAppend (conversion on the device)
        {
            std::ifstream probe(dir + "/raw_meta.bin");
            if (probe) {
                const auto rm = read_bin<double>(dir + "/raw_meta.bin"); // stamp, nearest, farthest, to_gps, ring lines
                LidarConverter ring(10.0, (int) rm[4], rm[1], rm[2]);
                std::vector<double> res;
                auto run = [&](auto fn) {
                    levins_host::check(levins_buffer_clear(levins_host::context()), "buffer_clear");
                    double start = 0, end = 0;
                    const size_t n = fn(start, end);
                    size_t in_ring = 0;
                    levins_host::check(levins_buffer_info(levins_host::context(), &in_ring, nullptr, nullptr), "buffer_info");
                    res.insert(res.end(), {(double) n, (double) in_ring, start, end});
                };
                const auto vel = read_bin<VelodynePoint>(dir + "/raw_velodyne.bin");
                run([&](double &s0, double &e0) { return ring.velodynePointCloudAppend(vel, rm[0], s0, e0, rm[3] != 0); });
                const auto ous = read_bin<OusterPoint>(dir + "/raw_ouster.bin");
                run([&](double &s0, double &e0) { return ring.ousterPointCloudAppend(ous, rm[0], s0, e0, rm[3] != 0); });
                const auto hes = read_bin<HesaiPoint>(dir + "/raw_hesai.bin");
                run([&](double &s0, double &e0) { return ring.hesaiPointCloudAppend(hes, rm[0], s0, e0, rm[3] != 0); });
                write_bin(dir + "/append_drivers.bin", res);
            }
        }
        levins_destroy(levins_host::context());
    } catch (const std::exception &e) {
        std::fprintf(stderr, "test_host failed: %s\n", e.what());
        return 1;
    }
    return 0;
}
