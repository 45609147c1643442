// ref_golden.cc -- drives the UNMODIFIED reference sources on the repo's golden fixtures and writes what they compute.
//
// TEST INFRASTRUCTURE (oracle side), not product code.  It cannot be built in the build container (no Eigen / PCL /
// FLANN / TBB / OpenCV / glog / yaml-cpp there); oracle/ref_harness/BUILD.md says where it can.  Nothing of the
// reference is copied: the reference's translation units are compiled from where they lie (LE_VINS_SRC) and this file only
// calls them:
//   stage 1   MISC::getPoseFromInsWindow (le_vins/common/misc.cc:62-78) + lidar::PointCloudCommon::relativeProjectionToCamera
//             (le_vins/lidar/pointcloud.cc:36-48) per point, in the arrangement of VINS::pointcloudStackForVisualTracking
//             (le_vins/le_vins/le_vins.cc:413-464; that member itself needs the whole estimator incl. Ceres, so only its
//             loop structure -- pop, num_projected, range gate -- is restated here, every arithmetic step is the reference's)
//   stage 2   visual::Tracking::depthAssociationProjection   (le_vins/visual/tracking.cc:298-368, private: reached with
//             the usual `#define private public` test hack, the header is not modified)
//   stage 3   visual::Tracking::depthAssociationPlaneFitting (tracking.cc:564-634) per query, plus one exact 5-NN query on
//             the reference's own pcl::KdTreeFLANN for the neighbour indices
//
//   ref_golden <fixture dir> <config.yaml>
// reads  points.bin (PointXYZIT) ins.bin (9 f64 per sample) queries.bin (3 f32) meta.bin (stamp, interval, td_b_l,
//        pose_b_c[12], pose_c_l[12], intrinsics fx fy cx cy)
// writes ref_cloud.bin ref_unit.bin ref_type.bin ref_xyzd.bin ref_knn_idx.bin ref_knn_d2.bin ref_summary.txt
#include <cstdio>
#include <deque>
#include <fstream>
#include <string>
#include <vector>

#define private public // visual::Tracking's depth-association members are private (tracking.h:92-95)
#include "visual/tracking.h"
#undef private

#include "common/misc.h"
#include "lidar/pointcloud.h"
#include "visual/visual_mappoint.h"

template <typename T> static std::vector<T> read_bin(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    const size_t bytes = f.tellg();
    f.seekg(0);
    std::vector<T> v(bytes / sizeof(T));
    f.read(reinterpret_cast<char *>(v.data()), bytes);
    return v;
}
template <typename T> static void write_bin(const std::string &path, const T *p, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(p), n * sizeof(T));
}

struct RawPoint { // byte image of PointXYZIT (le_vins/lidar/lidar.h:28-33)
    float x, y, z, pad0, intensity, pad1;
    double time;
};
struct OutPoint { // byte image of pcl::PointXYZI
    float x, y, z, w, intensity, pad[3];
};

int main(int argc, char **argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: ref_golden <fixture dir> <config.yaml>\n");
        return 2;
    }
    const std::string dir = argv[1], config = argv[2];
    const auto raw = read_bin<RawPoint>(dir + "/points.bin");
    const auto insv = read_bin<double>(dir + "/ins.bin");
    const auto queries = read_bin<float>(dir + "/queries.bin");
    const auto meta = read_bin<double>(dir + "/meta.bin");
    const double stamp = meta[0], interval = meta[1], td_b_l = meta[2];
    Pose pose_b_c, pose_c_l;
    std::memcpy(pose_b_c.R.data(), &meta[3], 72);
    std::memcpy(pose_b_c.t.data(), &meta[12], 24);
    std::memcpy(pose_c_l.R.data(), &meta[15], 72);
    std::memcpy(pose_c_l.t.data(), &meta[24], 24);

    std::deque<std::pair<IMU, IntegrationState>> ins_window;
    for (size_t i = 0; i + 8 < insv.size(); i += 9) {
        IMU imu;
        imu.time = insv[i];
        IntegrationState st;
        st.time = insv[i + 1];
        st.p = Eigen::Vector3d(insv[i + 2], insv[i + 3], insv[i + 4]);
        st.q = Eigen::Quaterniond(insv[i + 8], insv[i + 5], insv[i + 6], insv[i + 7]); // (w, x, y, z)
        ins_window.emplace_back(imu, st);
    }

    // ---- stage 1: le_vins.cc:413-464 ----
    size_t first = 0;
    while (first < raw.size() && raw[first].time < stamp - interval) first++; // :413-424
    size_t num_projected = raw.size() - first;
    for (size_t k = first; k < raw.size(); k++)
        if (raw[k].time > stamp) { // :431-437
            num_projected = k - first;
            break;
        }
    PointCloud pointcloud;
    pointcloud.resize(num_projected); // :438
    Pose last_keyframe_pose;
    MISC::getPoseFromInsWindow(ins_window, pose_b_c, stamp, last_keyframe_pose); // :441-442
    for (size_t k = 0; k < num_projected; k++) {
        const RawPoint &r = raw[first + k];
        const float norm = Eigen::Vector3f(r.x, r.y, r.z).norm();
        if ((norm > visual::VisualMapPoint::FARTHEST_DEPTH) || (norm < visual::VisualMapPoint::NEAREST_DEPTH)) continue; // :447-451
        Pose pose;
        MISC::getPoseFromInsWindow(ins_window, pose_b_c, r.time + td_b_l, pose); // :455
        PointType point;
        point.getVector3fMap() = Eigen::Vector3f(r.x, r.y, r.z);
        pointcloud[k] = lidar::PointCloudCommon::relativeProjectionToCamera(pose_c_l, last_keyframe_pose, pose, point); // :461
    }
    static_assert(sizeof(PointType) == sizeof(OutPoint), "pcl::PointXYZI layout");
    write_bin(dir + "/ref_cloud.bin", reinterpret_cast<const OutPoint *>(pointcloud.points.data()), pointcloud.size());

    // ---- stages 2 and 3 on the reference's own Tracking object ----
    auto camera = visual::Camera::createCamera({meta[27], meta[28], meta[29], meta[30]}, {0, 0, 0, 0}, {1440, 1080});
    auto map = std::make_shared<visual::VisualMap>(10);
    visual::Tracking tracking(camera, map, nullptr, config, dir);
    const bool ok = tracking.depthAssociationProjection(pointcloud);
    size_t m = 0;
    if (ok) {
        m = tracking.unit_sphere_pointcloud_->size();
        write_bin(dir + "/ref_unit.bin", reinterpret_cast<const OutPoint *>(tracking.unit_sphere_pointcloud_->points.data()), m);
    }
    const size_t f = queries.size() / 3;
    std::vector<int8_t> type(f, -1);
    std::vector<float> xyzd(4 * f, 0.f), d2(5 * f, 0.f);
    std::vector<int32_t> idx(5 * f, -1);
    for (size_t k = 0; ok && k < f; k++) {
        PointType q;
        q.x = queries[3 * k];
        q.y = queries[3 * k + 1];
        q.z = queries[3 * k + 2];
        q.intensity = 0;
        std::vector<int> ki;
        std::vector<float> kd;
        if (m >= 5) {
            tracking.kdtree_->nearestKSearch(q, 5, ki, kd); // the same query depthAssociationPlaneFitting issues (:571)
            for (size_t j = 0; j < ki.size() && j < 5; j++) {
                idx[5 * k + j] = ki[j];
                d2[5 * k + j] = kd[j];
            }
            type[k] = static_cast<int8_t>(tracking.depthAssociationPlaneFitting(q));
        }
        xyzd[4 * k] = q.x;
        xyzd[4 * k + 1] = q.y;
        xyzd[4 * k + 2] = q.z;
        xyzd[4 * k + 3] = type[k] == 1 ? q.intensity : 0.f;
    }
    write_bin(dir + "/ref_type.bin", type.data(), f);
    write_bin(dir + "/ref_xyzd.bin", xyzd.data(), 4 * f);
    write_bin(dir + "/ref_knn_idx.bin", idx.data(), 5 * f);
    write_bin(dir + "/ref_knn_d2.bin", d2.data(), 5 * f);
    std::ofstream s(dir + "/ref_summary.txt");
    s << "first " << first << "\nn " << num_projected << "\nm " << m << "\n";
    std::printf("%s: first=%zu n=%zu M=%zu F=%zu\n", dir.c_str(), first, num_projected, m, f);
    return 0;
}
