// Stand-in for <pcl/point_types.h>: just enough of pcl::PointXYZ (16 bytes, x y z + padding) to compile and run the
// PLANE_RANSAC_WITH_PCL overloads of include/PlaneDetectRansac.h where PCL is not installed.  Test infrastructure.
#pragma once
namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f, data_w = 1.f;
};
}  // namespace pcl
