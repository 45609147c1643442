// MISC pose lookup of the reference (le_vins/common/misc.h:35-41): same names, signatures and return conventions, bodies in
// liblevins_b200.so.  These are the estimator's single-shot host calls (e.g. le_vins.cc:364); the library evaluates them with
// the same source its kernels use for the per-point case (csrc/device_math.cuh: interpolate_state / pose_with_extrinsic,
// compiled for host and device), so the frame pose the estimator sees and the keyframe pose inside k_prepare are one
// function.  Nothing is computed in this header.
#ifndef LEVINS_HOST_MISC_H
#define LEVINS_HOST_MISC_H

#include <cstring>
#include <deque>
#include <utility>
#include <vector>

#include "common/types.h"
#include "levins/abi.h"

namespace levins_host {

inline void to_abi(const Pose &pose, levins_pose &out) {
    std::memcpy(out.R, pose.R.data(), sizeof(out.R));
    std::memcpy(out.t, pose.t.data(), sizeof(out.t));
}
inline Pose from_abi(const levins_pose &in) {
    Pose pose;
    std::memcpy(pose.R.data(), in.R, sizeof(in.R));
    std::memcpy(pose.t.data(), in.t, sizeof(in.t));
    return pose;
}
inline void to_abi(const IMU &imu, const IntegrationState &state, levins_ins_sample &s) {
    s.time = imu.time;
    s.state_time = state.time;
    std::memcpy(s.p, state.p.data(), sizeof(s.p));
    std::memcpy(s.q, state.q.coeffs(), sizeof(s.q));
}
inline void to_abi(const std::deque<std::pair<IMU, IntegrationState>> &window, std::vector<levins_ins_sample> &out) {
    out.resize(window.size());
    for (size_t i = 0; i < window.size(); i++) to_abi(window[i].first, window[i].second, out[i]);
}

} // namespace levins_host

class MISC {
public:
    typedef std::deque<std::pair<IMU, IntegrationState>> InsWindow;

    static size_t getInsWindowIndex(const InsWindow &window, double time) {
        std::vector<levins_ins_sample> w;
        levins_host::to_abi(window, w);
        return levins_ins_window_index(w.data(), w.size(), time);
    }

    static void statePoseInterpolation(const IntegrationState &state0, const IntegrationState &state1, double midtime,
                                       IntegrationState &state) {
        levins_ins_sample s0, s1, out;
        levins_host::to_abi(IMU(), state0, s0);
        levins_host::to_abi(IMU(), state1, s1);
        levins_state_interpolate(&s0, &s1, midtime, &out);
        state.time = out.state_time;
        state.p = Vector3d(out.p[0], out.p[1], out.p[2]);
        state.q = Quaterniond(out.q);
    }

    static Pose stateToPoseWithExtrinsic(const IntegrationState &state, const Pose &pose_ext) {
        levins_ins_sample s;
        levins_pose ext, out;
        levins_host::to_abi(IMU(), state, s);
        levins_host::to_abi(pose_ext, ext);
        levins_state_to_pose(&s, &ext, &out);
        return levins_host::from_abi(out);
    }

    static bool getPoseFromInsWindow(const InsWindow &window, const Pose &pose_ext, double time, Pose &pose) {
        std::vector<levins_ins_sample> w;
        levins_host::to_abi(window, w);
        levins_pose ext, out;
        levins_host::to_abi(pose_ext, ext);
        const int interpolated = levins_pose_at(w.data(), w.size(), &ext, time, &out);
        pose = levins_host::from_abi(out);
        return interpolated != 0;
    }
};

#endif
