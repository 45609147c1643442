// The Camera functions on the path (le_vins/visual/camera.cc:72-74, 123-127, 133-135, 149-151).  The per-feature use inside
// the association drivers runs on the device (levins_associate_pixels); the three one-liners stay for single host calls.
#ifndef LEVINS_HOST_CAMERA_H
#define LEVINS_HOST_CAMERA_H

#include <vector>

#include "common/types.h"
#include "levins/abi.h"

namespace visual {

struct Point2f {
    float x = 0, y = 0;
};

class Camera {
public:
    Camera(double fx, double fy, double cx, double cy, double skew = 0, const std::vector<double> &distortion = {})
        : fx_(fx), fy_(fy), cx_(cx), cy_(cy), skew_(skew) {
        for (size_t i = 0; i < distortion.size() && i < 5; i++) d_[i] = distortion[i]; // k1 k2 p1 p2 [k3] (camera.cc:60-67)
    }
    // what the device-side feature step needs of this camera
    levins_camera abi() const { return {fx_, fy_, cx_, cy_, skew_, d_[0], d_[1], d_[2], d_[3], d_[4]}; }
    Vector3d pixel2cam(const Point2f &pixel) const {
        const double y = (pixel.y - cy_) / fy_;
        const double x = (pixel.x - cx_ - skew_ * y) / fx_;
        return {x, y, 1.0};
    }
    Vector3d pixel2unitcam(const Point2f &pixel) const { return pixel2cam(pixel).normalized(); }
    static Vector3d cam2world(const Vector3d &cam, const Pose &pose) { return pose.R * cam + pose.t; }

private:
    double fx_, fy_, cx_, cy_, skew_;
    double d_[5] = {0, 0, 0, 0, 0};
};

} // namespace visual
#endif
