// lidar/lidar.h of the reference (le_vins/lidar/lidar.h:28-33, 85-92): point records and typedefs, same names, same
// 32-byte layouts (they are memcpy'd to the device as they are).
#ifndef LEVINS_HOST_LIDAR_H
#define LEVINS_HOST_LIDAR_H

#include <vector>

#include "common/eigen_lite.h"

#if !defined(LEVINS_USE_EIGEN_PCL)
// PointXYZIT: PCL_ADD_POINT4D + intensity + double time, EIGEN_ALIGN16
struct alignas(16) PointXYZIT {
    float x = 0, y = 0, z = 0, data3 = 1.0f;
    float intensity = 0;
    float pad_ = 0;
    double time = 0;
};
#else
struct PointXYZIT {
    PCL_ADD_POINT4D;
    float intensity;
    double time;
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
} EIGEN_ALIGN16;
#endif
static_assert(sizeof(PointXYZIT) == 32, "PointXYZIT must be 32 bytes (le_vins/lidar/lidar.h:28-33)");

enum LidarType { Livox = 1, Velodyne = 2, Ouster = 3, Hesai = 4 };

typedef PointXYZIT PointTypeCustom;
typedef pcl::PointCloud<PointTypeCustom> PointCloudCustom;
typedef pcl::PointCloud<PointTypeCustom>::Ptr PointCloudCustomPtr;

typedef pcl::PointXYZI PointType;
typedef pcl::PointCloud<PointType> PointCloud;
typedef pcl::PointCloud<PointType>::Ptr PointCloudPtr;
typedef std::vector<PointType> PointVector;
static_assert(sizeof(PointType) == 32, "pcl::PointXYZI must be 32 bytes");

#endif
