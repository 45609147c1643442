// CPU-only driver for the part of the host mirror that needs no GPU: MISC pose lookup (host exports of the library), the single-point
// PointCloudCommon statics, Camera.  tests/test_host_mirror.py feeds it binary vectors
// and compares every output with the oracle.
//   test_host_cpu <dir>   reads ins.bin times.bin ext.bin pts.bin   writes poses.bin proj.bin cam.bin
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "common/misc.h"
#include "lidar/pointcloud.h"
#include "visual/camera.h"
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
template <typename T> static void write_bin(const std::string &path, const std::vector<T> &v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), v.size() * sizeof(T));
}
static void push_pose(std::vector<double> &out, const Pose &p) {
    out.insert(out.end(), p.R.data(), p.R.data() + 9);
    out.insert(out.end(), p.t.data(), p.t.data() + 3);
}
static Pose pose_from(const double *m) {
    Pose p;
    std::memcpy(p.R.data(), m, 72);
    std::memcpy(p.t.data(), m + 9, 24);
    return p;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    const auto ins_raw = read_bin<double>(dir + "/ins.bin");
    const auto times = read_bin<double>(dir + "/times.bin");
    const auto ext = read_bin<double>(dir + "/ext.bin"); // pose_b_c[12], pose_c_l[12], intr[5]
    const auto pts = read_bin<PointType>(dir + "/pts.bin");
    MISC::InsWindow window;
    for (size_t i = 0; i + 8 < ins_raw.size(); i += 9) {
        IMU imu;
        imu.time = ins_raw[i];
        IntegrationState st;
        st.time = ins_raw[i + 1];
        st.p = Vector3d(ins_raw[i + 2], ins_raw[i + 3], ins_raw[i + 4]);
        st.q = Quaterniond(&ins_raw[i + 5]);
        window.push_back({imu, st});
    }
    const Pose pose_b_c = pose_from(&ext[0]), pose_c_l = pose_from(&ext[12]);

    // MISC::getInsWindowIndex / getPoseFromInsWindow
    std::vector<double> poses;
    std::vector<Pose> pose_list;
    for (double t : times) {
        Pose pose;
        const bool ok = MISC::getPoseFromInsWindow(window, pose_b_c, t, pose);
        poses.push_back((double) MISC::getInsWindowIndex(window, t));
        poses.push_back(ok ? 1.0 : 0.0);
        push_pose(poses, pose);
        pose_list.push_back(pose);
    }
    write_bin(dir + "/poses.bin", poses);

    // PointCloudCommon single-point statics
    std::vector<PointType> proj;
    for (size_t k = 0; k < pts.size(); k++) {
        const Pose &p0 = pose_list[k % pose_list.size()], &p1 = pose_list[(k * 7 + 3) % pose_list.size()];
        proj.push_back(lidar::PointCloudCommon::pointProjection(pose_c_l, pts[k]));
        proj.push_back(lidar::PointCloudCommon::relativeProjection(p0, p1, pts[k]));
        proj.push_back(lidar::PointCloudCommon::relativeProjectionToCamera(pose_c_l, p0, p1, pts[k]));
    }
    write_bin(dir + "/proj.bin", proj);

    // Camera + isGoodDepth
    visual::Camera camera(ext[24], ext[25], ext[26], ext[27], ext[28]);
    std::vector<double> cam;
    for (size_t k = 0; k < pts.size(); k++) {
        const visual::Point2f px{pts[k].x * 10.0f + 600.0f, pts[k].y * 10.0f + 500.0f};
        const Vector3d pc = camera.pixel2cam(px), puc = camera.pixel2unitcam(px);
        const Vector3d pw = visual::Camera::cam2world(puc, pose_list[k % pose_list.size()]);
        for (int i = 0; i < 3; i++) cam.push_back(pc[i]);
        for (int i = 0; i < 3; i++) cam.push_back(puc[i]);
        for (int i = 0; i < 3; i++) cam.push_back(pw[i]);
        cam.push_back(visual::VisualMapPoint::isGoodDepth(pts[k].z, 1.0) ? 1.0 : 0.0);
    }
    write_bin(dir + "/cam.bin", cam);

    return 0;
}
