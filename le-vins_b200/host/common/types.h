// common/types.h, common/integration_state.h of the reference, reduced to what the LiDAR depth path touches
// (le_vins/common/types.h:41-54, le_vins/common/integration_state.h:33-49).
#ifndef LEVINS_HOST_TYPES_H
#define LEVINS_HOST_TYPES_H

#include "common/eigen_lite.h"

using Eigen::Matrix3d;
using Eigen::Quaterniond;
using Eigen::Vector3d;

typedef struct IMU {
    double time = 0;
    double dt = 0;
    Vector3d dtheta;
    Vector3d dvel;
    double odovel = 0;
} IMU;

typedef struct Pose {
    Matrix3d R;
    Vector3d t;
} Pose;

// only time / p / q are read on this path (misc.cc:84-94)
typedef struct IntegrationState {
    double time = 0;
    Vector3d p{0, 0, 0};
    Quaterniond q{0, 0, 0, 0};
} IntegrationState;

#endif
