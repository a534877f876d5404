// Drives the PLANE_RANSAC_WITH_PCL overload of include/PlaneDetectRansac.h — the surface BASELINE's north_star names:
// (pcl::PointCloud<pcl::PointXYZ>::Ptr, threshold, max iterations, min plane size) -> pcl::ModelCoefficients +
// pcl::PointIndices per plane — over the stand-in PCL headers in tests/pcl_stub (PCL itself is not in this image).
//   pcl_overload_check <cloud.f32> <threshold> <max_iterations> <min_plane_size> [pin]
#define PLANE_RANSAC_WITH_PCL
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "PlaneDetectRansac.h"

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  if (!f) return 2;
  pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>());
  float v[3];
  while (f.read(reinterpret_cast<char*>(v), sizeof(v))) {
    pcl::PointXYZ p;
    p.x = v[0]; p.y = v[1]; p.z = v[2];
    cloud->points.push_back(p);
  }
  cloud->width = (uint32_t)cloud->points.size();
  cloud->height = 1;
  const size_t n = cloud->points.size();
  plane_detect_ransac::PlaneDetectRansac det(0);
  if (!det.ok()) { std::fprintf(stderr, "%s\n", det.last_error().c_str()); return 1; }
  det.setDistanceThreshold(std::atof(argv[2]));
  det.setMaxIterations(std::atoi(argv[3]));
  det.setMinPlaneSize(std::atoi(argv[4]));
  const bool pin = argc > 5 && std::string(argv[5]) == "pin";
  void* pinned_at = cloud->points.data();
  if (pin && !det.pinCallerMemory(pinned_at, n * sizeof(pcl::PointXYZ))) { std::fprintf(stderr, "%s\n", det.last_error().c_str()); return 1; }
  std::vector<pcl::ModelCoefficients> coefficients;
  std::vector<pcl::PointIndices> inliers;
  if (!det.detect(cloud, coefficients, inliers)) { std::fprintf(stderr, "%s\n", det.last_error().c_str()); return 1; }
  if (pin) det.unpinCallerMemory(pinned_at);
  std::printf("points %zu planes %zu remaining %zu width %u\n", n, coefficients.size(), cloud->points.size(), cloud->width);
  for (size_t k = 0; k < coefficients.size(); ++k) {
    unsigned long long h = 0;
    for (size_t i = 0; i < inliers[k].indices.size(); ++i) h += ((unsigned long long)(unsigned)inliers[k].indices[i] + 1ull) * ((unsigned long long)i + 1ull);
    std::printf("plane %zu inliers %zu coeff %a %a %a %a hash %llu\n", k, inliers[k].indices.size(), coefficients[k].values[0],
                coefficients[k].values[1], coefficients[k].values[2], coefficients[k].values[3], h);
  }
  double sx = 0;
  for (const pcl::PointXYZ& p : cloud->points) sx += p.x + p.y + p.z;
  std::printf("remaining_sum %a\n", sx);
  return 0;
}
