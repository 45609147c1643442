// The three Camera functions on the path (le_vins/visual/camera.cc:123-127, 133-135, 149-151).
#ifndef LEVINS_HOST_CAMERA_H
#define LEVINS_HOST_CAMERA_H

#include "common/types.h"

namespace visual {

struct Point2f {
    float x = 0, y = 0;
};

class Camera {
public:
    Camera(double fx, double fy, double cx, double cy, double skew = 0) : fx_(fx), fy_(fy), cx_(cx), cy_(cy), skew_(skew) {}
    Vector3d pixel2cam(const Point2f &pixel) const {
        const double y = (pixel.y - cy_) / fy_;
        const double x = (pixel.x - cx_ - skew_ * y) / fx_;
        return {x, y, 1.0};
    }
    Vector3d pixel2unitcam(const Point2f &pixel) const { return pixel2cam(pixel).normalized(); }
    static Vector3d cam2world(const Vector3d &cam, const Pose &pose) { return pose.R * cam + pose.t; }

private:
    double fx_, fy_, cx_, cy_, skew_;
};

} // namespace visual
#endif
