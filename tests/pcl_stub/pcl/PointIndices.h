// Stand-in for <pcl/PointIndices.h> (see point_types.h).
#pragma once
#include <vector>
namespace pcl {
struct PointIndices {
  std::vector<int> indices;
};
}  // namespace pcl
