// eigen_lite.h -- the sliver of Eigen / PCL the reference's surface for this path needs, for builds where the real
// libraries are absent (this image has neither).  With Eigen and PCL installed, define LEVINS_USE_EIGEN_PCL and the
// real types are used instead (layouts are identical: Matrix3d column-major, Quaterniond coeffs x,y,z,w,
// pcl::PointXYZI 32 B).
#ifndef LEVINS_HOST_EIGEN_LITE_H
#define LEVINS_HOST_EIGEN_LITE_H

#if defined(LEVINS_USE_EIGEN_PCL)
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else

#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

namespace Eigen {

struct Vector3d {
    double v[3]{0, 0, 0};
    Vector3d() = default;
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double &x() { return v[0]; }
    double &y() { return v[1]; }
    double &z() { return v[2]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    const double *data() const { return v; }
    double *data() { return v; }
    Vector3d operator+(const Vector3d &o) const { return {v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]}; }
    Vector3d operator-(const Vector3d &o) const { return {v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}; }
    Vector3d operator*(double s) const { return {v[0] * s, v[1] * s, v[2] * s}; }
    double dot(const Vector3d &o) const { return (v[0] * o.v[0] + v[1] * o.v[1]) + v[2] * o.v[2]; }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    Vector3d normalized() const {
        const double z = squaredNorm();
        if (z > 0) {
            const double s = std::sqrt(z);
            return {v[0] / s, v[1] / s, v[2] / s};
        }
        return *this;
    }
};

// column-major 3x3, like Eigen::Matrix3d
struct Matrix3d {
    double m[9]{1, 0, 0, 0, 1, 0, 0, 0, 1};
    double &operator()(int r, int c) { return m[3 * c + r]; }
    double operator()(int r, int c) const { return m[3 * c + r]; }
    const double *data() const { return m; }
    double *data() { return m; }
    static Matrix3d Identity() { return Matrix3d(); }
    Matrix3d transpose() const {
        Matrix3d t;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) t(r, c) = (*this)(c, r);
        return t;
    }
    Vector3d operator*(const Vector3d &x) const {
        Vector3d y;
        for (int i = 0; i < 3; i++) y[i] = (m[i] * x[0] + m[3 + i] * x[1]) + m[6 + i] * x[2];
        return y;
    }
    Matrix3d operator*(const Matrix3d &b) const {
        Matrix3d c;
        for (int j = 0; j < 3; j++)
            for (int i = 0; i < 3; i++) c.m[3 * j + i] = (m[i] * b.m[3 * j] + m[3 + i] * b.m[3 * j + 1]) + m[6 + i] * b.m[3 * j + 2];
        return c;
    }
};

// coefficient order x, y, z, w like Eigen::Quaterniond::coeffs()
struct Quaterniond {
    double c[4]{0, 0, 0, 1};
    Quaterniond() = default;
    Quaterniond(double w, double x, double y, double z) : c{x, y, z, w} {}
    explicit Quaterniond(const double *xyzw) { std::memcpy(c, xyzw, sizeof(c)); }
    double x() const { return c[0]; }
    double y() const { return c[1]; }
    double z() const { return c[2]; }
    double w() const { return c[3]; }
    const double *coeffs() const { return c; }
    double squaredNorm() const { return ((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]) + c[3] * c[3]; }
    Quaterniond inverse() const {
        const double n2 = squaredNorm();
        if (n2 > 0) return Quaterniond(c[3] / n2, -c[0] / n2, -c[1] / n2, -c[2] / n2);
        return Quaterniond(0, 0, 0, 0);
    }
    void normalize() {
        const double z = squaredNorm();
        if (z > 0) {
            const double s = std::sqrt(z);
            for (double &v : c) v /= s;
        }
    }
    Quaterniond operator*(const Quaterniond &b) const {
        const double ax = c[0], ay = c[1], az = c[2], aw = c[3];
        const double bx = b.c[0], by = b.c[1], bz = b.c[2], bw = b.c[3];
        return Quaterniond(aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                           aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx);
    }
    Matrix3d toRotationMatrix() const {
        const double x = c[0], y = c[1], z = c[2], w = c[3];
        const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
        const double twx = tx * w, twy = ty * w, twz = tz * w;
        const double txx = tx * x, txy = ty * x, txz = tz * x;
        const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
        Matrix3d R;
        R(0, 0) = 1 - (tyy + tzz);
        R(0, 1) = txy - twz;
        R(0, 2) = txz + twy;
        R(1, 0) = txy + twz;
        R(1, 1) = 1 - (txx + tzz);
        R(1, 2) = tyz - twx;
        R(2, 0) = txz - twy;
        R(2, 1) = tyz + twx;
        R(2, 2) = 1 - (txx + tyy);
        return R;
    }
};

} // namespace Eigen

namespace pcl {

// pcl::PointXYZI: 32 B, data[3] = 1.0f
struct alignas(16) PointXYZI {
    float x = 0, y = 0, z = 0, data3 = 1.0f;
    float intensity = 0, pad_[3] = {0, 0, 0};
};

template <typename PointT> class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    std::vector<PointT> points;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); }
    void resize(size_t n) { points.resize(n); }
    void reserve(size_t n) { points.reserve(n); }
    void push_back(const PointT &p) { points.push_back(p); }
    PointT &operator[](size_t i) { return points[i]; }
    const PointT &operator[](size_t i) const { return points[i]; }
};

} // namespace pcl

#endif // LEVINS_USE_EIGEN_PCL
#endif
