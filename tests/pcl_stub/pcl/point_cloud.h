// Stand-in for <pcl/point_cloud.h> (see point_types.h): points / width / height and the Ptr typedef.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;  // boost::shared_ptr in PCL 1.8: same operator->
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  size_t size() const { return points.size(); }
};
}  // namespace pcl
