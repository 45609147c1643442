// lidar::PointCloudCommon of the reference (le_vins/lidar/pointcloud.h:29-52): same statics.  The single-point
// projections are trivial host inlines (they exist for API compatibility; the per-point loop that called them at
// le_vins.cc:461 is k_deskew_bin); the whole-cloud projection (pointcloud.cc:50-63, tbb::parallel_for) runs on the GPU.
#ifndef LEVINS_HOST_POINTCLOUD_H
#define LEVINS_HOST_POINTCLOUD_H

#include <cstring>

#include <vector>

#include "common/device.h"
#include "common/small_linalg.h"
#include "common/types.h"
#include "lidar/lidar.h"

namespace lidar {

class PointCloudCommon {
public:
    static PointType pointProjection(const Pose &pose, const PointType &point) { // pointcloud.cc:65-71
        PointType projected = point;
        const Vector3d r = pose.R * Vector3d(point.x, point.y, point.z) + pose.t;
        projected.x = static_cast<float>(r[0]);
        projected.y = static_cast<float>(r[1]);
        projected.z = static_cast<float>(r[2]);
        return projected;
    }

    static PointType relativeProjection(const Pose &pose0, const Pose &pose1, const PointType &point_l1) { // :28-34
        Pose rel;
        rel.R = pose0.R.transpose() * pose1.R;
        rel.t = pose0.R.transpose() * (pose1.t - pose0.t);
        return pointProjection(rel, point_l1);
    }

    static PointType relativeProjectionToCamera(const Pose &pose_c_l, const Pose &pose_cam0, const Pose &pose_cam1,
                                                const PointType &point_lidar) { // :36-48
        const PointType point_cam1 = pointProjection(pose_c_l, point_lidar);
        Pose rel;
        rel.R = pose_cam0.R.transpose() * pose_cam1.R;
        rel.t = pose_cam0.R.transpose() * (pose_cam1.t - pose_cam0.t);
        return pointProjection(rel, point_cam1);
    }

    // GPU: levins_point_cloud_projection (k_project)
    static void pointCloudProjection(const Pose &pose, PointCloudPtr src, PointCloudPtr dst) { // :50-63
        dst->resize(src->size());
        levins_pose p;
        std::memcpy(p.R, pose.R.data(), sizeof(p.R));
        std::memcpy(p.t, pose.t.data(), sizeof(p.t));
        levins_host::check(levins_point_cloud_projection(levins_host::context(), &p,
                                                         reinterpret_cast<const levins_point_xyzi *>(src->points.data()),
                                                         reinterpret_cast<levins_point_xyzi *>(dst->points.data()), src->size()),
                           "pointCloudProjection");
    }

    static double pointSquareDistance(const PointType &point0, const PointType &point1) { // :107-109
        const float dx = point0.x - point1.x, dy = point0.y - point1.y, dz = point0.z - point1.z;
        return dx * dx + (dy * dy + dz * dz);
    }

    // pointcloud.cc:73-105 (no call site in the reference; kept for the drop-in surface)
    static bool planeEstimation(const PointVector &points, double threshold, Vector3d &unit_normal_vector, double &norm_inverse,
                                std::vector<double> &distance) {
        if (points.size() < (size_t) PLANE_ESTIMATION_POINT_SZIE) return false;
        double A[15], b[5], n[3];
        for (int k = 0; k < PLANE_ESTIMATION_POINT_SZIE; k++) {
            A[3 * k] = points[k].x;
            A[3 * k + 1] = points[k].y;
            A[3 * k + 2] = points[k].z;
            b[k] = -1.0;
        }
        levins_host::colpiv_qr_solve_5x3(A, b, n); // A.colPivHouseholderQr().solve(b)
        const Vector3d normal_vector(n[0], n[1], n[2]);
        norm_inverse = 1.0 / normal_vector.norm();
        unit_normal_vector = normal_vector * norm_inverse;
        distance.clear();
        for (int k = 0; k < PLANE_ESTIMATION_POINT_SZIE; k++) {
            const double dis = unit_normal_vector.dot(Vector3d(points[k].x, points[k].y, points[k].z)) + norm_inverse;
            distance.push_back(dis);
            if (std::fabs(dis) > threshold) return false;
        }
        return true;
    }

    // pointcloud.cc:111-190 (no call site in the reference): point-to-plane sigma by first-order covariance propagation
    // through the eigen-decomposition of the neighbourhood covariance
    static double pointToPlaneStd(const PointVector &nearest, const PointType &target, double point_std) {
        const double pvv = point_std * point_std;
        const size_t point_size = nearest.size();
        const double inv = 1.0 / static_cast<double>(point_size);
        double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, center[3] = {0, 0, 0};
        for (const auto &p : nearest) {
            const double pv[3] = {p.x, p.y, p.z};
            for (int i = 0; i < 3; i++) {
                center[i] += pv[i];
                for (int j = 0; j < 3; j++) cov[3 * i + j] += pv[i] * pv[j];
            }
        }
        for (int i = 0; i < 3; i++) center[i] *= inv;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) cov[3 * i + j] = cov[3 * i + j] * inv - center[i] * center[j];
        double evals[3], evecs[9]; // evecs column m = evecs[3*i + m]
        levins_host::sym_eigen_3x3(cov, evals, evecs);
        int emin = 0;
        for (int m = 1; m < 3; m++)
            if (evals[m] < evals[emin]) emin = m;
        const double nrm[3] = {evecs[emin], evecs[3 + emin], evecs[6 + emin]};
        double plane_cov[6][6] = {};
        for (size_t i = 0; i < point_size; i++) {
            const double d[3] = {nearest[i].x - center[0], nearest[i].y - center[1], nearest[i].z - center[2]};
            double F[3][3] = {};
            for (int m = 0; m < 3; m++) {
                if (m == emin) continue;
                const double em[3] = {evecs[m], evecs[3 + m], evecs[6 + m]};
                const double scale = inv / (evals[emin] - evals[m]);
                // row m = scale * d^T (e_m n^T + n e_m^T)
                const double de = d[0] * em[0] + d[1] * em[1] + d[2] * em[2], dn = d[0] * nrm[0] + d[1] * nrm[1] + d[2] * nrm[2];
                for (int c = 0; c < 3; c++) F[m][c] = scale * (de * nrm[c] + dn * em[c]);
            }
            double J[6][3] = {};
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    double s = 0;
                    for (int m = 0; m < 3; m++) s += evecs[3 * r + m] * F[m][c];
                    J[r][c] = s;
                }
            for (int r = 0; r < 3; r++) J[3 + r][r] = inv;
            for (int r = 0; r < 6; r++)
                for (int c = 0; c < 6; c++) {
                    double s = 0;
                    for (int k = 0; k < 3; k++) s += J[r][k] * J[c][k];
                    plane_cov[r][c] += pvv * s;
                }
        }
        const double jac[6] = {target.x - center[0], target.y - center[1], target.z - center[2], -nrm[0], -nrm[1], -nrm[2]};
        double var = 0;
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) var += jac[r] * plane_cov[r][c] * jac[c];
        var += pvv * (nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
        return std::sqrt(var);
    }

public:
    static constexpr int PLANE_ESTIMATION_POINT_SZIE = 5;
};

} // namespace lidar
#endif
