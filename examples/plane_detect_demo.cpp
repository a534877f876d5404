// plane_detect_demo — drives the C++ shim (include/PlaneDetectRansac.h) the way Dialog's
// on_autoPerformAction_triggered drives PlaneDetect.h (Dialog/PCLViewer.cpp:1180-1235): load a cloud,
// run plane detection with the config.txt parameters, print one line per plane and the points left.
//
//   plane_detect_demo <cloud.pcd | cloud.f32> [distance_threshold] [max_iterations] [min_plane_size]
//
// .pcd: ASCII PCD v0.7 with x y z as the first three fields (the format of Dialog/double_shadow.pcd);
// .f32: raw little-endian float32 x,y,z triples.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "PlaneDetectRansac.h"

using plane_detect_ransac::PlaneDetectRansac;
using plane_detect_ransac::PlaneRecord;
using plane_detect_ransac::PointXYZ;

static bool load_cloud(const std::string& path, std::vector<PointXYZ>& out) {
  if (path.size() > 4 && path.substr(path.size() - 4) == ".pcd") {
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    bool data = false;
    while (std::getline(f, line)) {
      if (!data) {
        if (line.rfind("DATA", 0) == 0) {
          if (line.find("ascii") == std::string::npos) return false;
          data = true;
        }
        continue;
      }
      std::istringstream ss(line);
      PointXYZ p{0, 0, 0, 1.0f};
      if (ss >> p.x >> p.y >> p.z) out.push_back(p);
    }
    return data;
  }
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  float v[3];
  while (f.read(reinterpret_cast<char*>(v), sizeof(v))) out.push_back(PointXYZ{v[0], v[1], v[2], 1.0f});
  return true;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <cloud.pcd|cloud.f32> [threshold] [max_iterations] [min_plane_size]\n", argv[0]);
    return 2;
  }
  std::vector<PointXYZ> cloud;
  if (!load_cloud(argv[1], cloud)) {
    std::fprintf(stderr, "cannot read %s\n", argv[1]);
    return 2;
  }
  PlaneDetectRansac det(0);
  if (!det.ok()) {
    std::fprintf(stderr, "%s\n", det.last_error().c_str());
    return 1;
  }
  if (argc > 2) det.setDistanceThreshold(std::atof(argv[2]));  // config.txt T_dist_point_plane
  if (argc > 3) det.setMaxIterations(std::atoi(argv[3]));
  if (argc > 4) det.setMinPlaneSize(std::atoi(argv[4]));      // config.txt T_num_of_single_plane
  const size_t n = cloud.size();
  std::vector<PlaneRecord> planes;
  if (!det.detect(cloud, planes)) {
    std::fprintf(stderr, "%s\n", det.last_error().c_str());
    return 1;
  }
  std::printf("points %zu planes %zu remaining %zu\n", n, planes.size(), cloud.size());
  for (size_t k = 0; k < planes.size(); ++k)
    std::printf("plane %zu inliers %zu coeff %a %a %a %a\n", k, planes[k].indices.size(), planes[k].coeff[0],
                planes[k].coeff[1], planes[k].coeff[2], planes[k].coeff[3]);
  return 0;
}
