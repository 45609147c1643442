// MISC pose lookup of the reference (le_vins/common/misc.h:35-41, misc.cc:25-103) and the two Rotation helpers it uses
// (rotation.cc:59-68).  Same names, signatures and semantics.  These are the estimator's single-shot host calls
// (e.g. le_vins.cc:364); the per-point use inside the deskew loop (le_vins.cc:455) is what k_deskew_bin replaces.
#ifndef LEVINS_HOST_MISC_H
#define LEVINS_HOST_MISC_H

#include <cfloat>
#include <cmath>
#include <deque>
#include <utility>

#include "common/types.h"

class Rotation {
public:
    static Quaterniond rotvec2quaternion(const Vector3d &rotvec) { // rotation.cc:59-63
        const double angle = rotvec.norm();
        const Vector3d vec = rotvec.normalized();
        const double ha = 0.5 * angle;
        const double s = std::sin(ha);
        return Quaterniond(std::cos(ha), s * vec[0], s * vec[1], s * vec[2]);
    }
    static Vector3d quaternion2vector(const Quaterniond &q) { // rotation.cc:65-68 (Eigen::AngleAxisd(q))
        Vector3d v(q.x(), q.y(), q.z());
        double n = v.norm();
        if (n < DBL_EPSILON) {
            const double m = std::fmax(std::fabs(v[0]), std::fmax(std::fabs(v[1]), std::fabs(v[2])));
            n = m > 0 ? m * Vector3d(v[0] / m, v[1] / m, v[2] / m).norm() : 0.0;
        }
        if (n != 0) {
            const double angle = 2 * std::atan2(n, std::fabs(q.w()));
            if (q.w() < 0) n = -n;
            return Vector3d(v[0] / n, v[1] / n, v[2] / n) * angle;
        }
        return Vector3d(0, 0, 0);
    }
};

class MISC {
public:
    typedef std::deque<std::pair<IMU, IntegrationState>> InsWindow;

    static size_t getInsWindowIndex(const InsWindow &window, double time) { // misc.cc:25-60
        if (window.empty() || (window.front().first.time > time) || (window.back().first.time <= time)) return 0;
        size_t index = 0, sta = 0, end = window.size();
        int counts = 0;
        while (true) {
            const size_t mid = (sta + end) / 2;
            const double first = window[mid - 1].first.time, second = window[mid].first.time;
            if ((first <= time) && (time < second)) {
                index = mid;
                break;
            } else if (first > time) {
                end = mid;
            } else if (second <= time) {
                sta = mid;
            }
            if (counts++ > 15) break;
        }
        return index;
    }

    static void statePoseInterpolation(const IntegrationState &state0, const IntegrationState &state1, double midtime,
                                       IntegrationState &state) { // misc.cc:80-95
        const Vector3d dp = state1.p - state0.p;
        Quaterniond dq = state1.q.inverse() * state0.q;
        Vector3d rvec = Rotation::quaternion2vector(dq);
        const double scale = (midtime - state0.time) / (state1.time - state0.time);
        rvec = rvec * scale;
        dq = Rotation::rotvec2quaternion(rvec);
        state.time = midtime;
        state.p = state0.p + dp * scale;
        state.q = state0.q * dq.inverse();
        state.q.normalize();
    }

    static Pose stateToPoseWithExtrinsic(const IntegrationState &state, const Pose &pose_ext) { // misc.cc:97-103
        Pose pose;
        pose.t = state.p + state.q.toRotationMatrix() * pose_ext.t;
        pose.R = state.q.toRotationMatrix() * pose_ext.R;
        return pose;
    }

    static bool getPoseFromInsWindow(const InsWindow &window, const Pose &pose_ext, double time, Pose &pose) { // misc.cc:62-78
        const size_t index = getInsWindowIndex(window, time);
        IntegrationState state;
        if (index > 0) {
            statePoseInterpolation(window[index - 1].second, window[index].second, time, state);
            pose = stateToPoseWithExtrinsic(state, pose_ext);
            return true;
        }
        pose = stateToPoseWithExtrinsic(window.back().second, pose_ext);
        return false;
    }
};

#endif
