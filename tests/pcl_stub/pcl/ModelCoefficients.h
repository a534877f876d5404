// Stand-in for <pcl/ModelCoefficients.h> (see point_types.h).
#pragma once
#include <vector>
namespace pcl {
struct ModelCoefficients {
  std::vector<float> values;
};
}  // namespace pcl
