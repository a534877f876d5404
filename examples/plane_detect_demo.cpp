// plane_detect_demo — drives the C++ shim (include/PlaneDetectRansac.h) the way Dialog's
// on_autoPerformAction_triggered / on_runAgainAction_triggered drive PlaneDetect.h (Dialog/PCLViewer.cpp:1120-1235):
// load a cloud, run plane detection with the config.txt parameters, optionally the postProcessPlanes steps
// (re-absorption against plane outlines, clusterFilt) and a second detection on what is left; one line per result.
//
//   plane_detect_demo <cloud.pcd | cloud.f32> [distance_threshold] [max_iterations] [min_plane_size] [options]
//     --prob P               SACSegmentation::setProbability (1 = score every hypothesis)
//     --max-planes N
//     --pipeline SEED RADIUS TNUM BORDERS_OUT POST_T
//                            after detect(): postProcess() with isPointInPoly's distance threshold POST_T against the bounding rectangle of each plane's inliers
//                            (written to BORDERS_OUT so that a checker can replay the pass), clusterFilter(RADIUS, TNUM),
//                            runAgain()
//     --bench STEPS          the cloud in a page-locked buffer; STEPS timed detectViews() calls, host cloud in ->
//                            coefficients + index lists out, wall clock per call
//
// .pcd: PCD v0.7 through plane_ransac_load_pcd (ascii like Dialog/double_shadow.pcd, binary, binary_compressed);
// .f32: raw little-endian float32 x,y,z triples.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "PlaneDetectRansac.h"

using plane_detect_ransac::PinnedCloud;
using plane_detect_ransac::PlaneDetectRansac;
using plane_detect_ransac::PlaneRecord;
using plane_detect_ransac::PlaneView;
using plane_detect_ransac::PointXYZ;

static bool load_cloud(const std::string& path, std::vector<PointXYZ>& out) {
  if (path.size() > 4 && path.substr(path.size() - 4) == ".pcd") {
    // the library's reader (pcl::io::loadPCDFile's formats: ascii, binary, binary_compressed)
    pr_point* pts = nullptr;
    size_t n = 0;
    if (plane_ransac_load_pcd(path.c_str(), &pts, &n) != PR_OK) {
      std::fprintf(stderr, "%s\n", plane_ransac_last_error());
      return false;
    }
    out.resize(n);
    if (n) std::memcpy(out.data(), pts, n * sizeof(PointXYZ));
    plane_ransac_host_free(pts);
    return true;
  }
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  f.seekg(0, std::ios::end);
  const size_t n = (size_t)f.tellg() / 12;
  f.seekg(0);
  std::vector<float> raw(3 * n);
  f.read(reinterpret_cast<char*>(raw.data()), (std::streamsize)(raw.size() * sizeof(float)));
  out.resize(n);
  for (size_t i = 0; i < n; ++i) out[i] = PointXYZ{raw[3 * i], raw[3 * i + 1], raw[3 * i + 2], 1.0f};
  return true;
}

// sum of (index + 1) * (position + 1) mod 2^64: cheap to reproduce with numpy
static unsigned long long list_hash(const int32_t* idx, size_t n) {
  unsigned long long h = 0;
  for (size_t i = 0; i < n; ++i) h += ((unsigned long long)(uint32_t)idx[i] + 1ull) * ((unsigned long long)i + 1ull);
  return h;
}

// Outline of a plane's inliers: the bounding rectangle in an orthonormal basis of the plane, 8 vertices per side.
static std::vector<PointXYZ> outline(const std::vector<PointXYZ>& cloud, const PlaneRecord& pl) {
  const double n[3] = {pl.coeff[0], pl.coeff[1], pl.coeff[2]};
  double a[3] = {1, 0, 0};
  if (std::fabs(n[0]) > 0.9) { a[0] = 0; a[1] = 1; }
  double u[3] = {n[1] * a[2] - n[2] * a[1], n[2] * a[0] - n[0] * a[2], n[0] * a[1] - n[1] * a[0]};
  const double ul = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  for (double& c : u) c /= ul;
  const double v[3] = {n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0]};
  double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
  for (int32_t i : pl.indices) {
    const PointXYZ& p = cloud[(size_t)i];
    const double s = p.x * u[0] + p.y * u[1] + p.z * u[2], t = p.x * v[0] + p.y * v[1] + p.z * v[2];
    lo[0] = std::min(lo[0], s); hi[0] = std::max(hi[0], s);
    lo[1] = std::min(lo[1], t); hi[1] = std::max(hi[1], t);
  }
  const double d = -(double)pl.coeff[3];  // points of the plane: s u + t v + d n
  std::vector<PointXYZ> out;
  const int per = 8;
  auto push = [&](double s, double t) {
    out.push_back(PointXYZ{(float)(s * u[0] + t * v[0] + d * n[0]), (float)(s * u[1] + t * v[1] + d * n[1]),
                           (float)(s * u[2] + t * v[2] + d * n[2]), 1.0f});
  };
  for (int k = 0; k < per; ++k) push(lo[0] + (hi[0] - lo[0]) * k / per, lo[1]);
  for (int k = 0; k < per; ++k) push(hi[0], lo[1] + (hi[1] - lo[1]) * k / per);
  for (int k = 0; k < per; ++k) push(hi[0] - (hi[0] - lo[0]) * k / per, hi[1]);
  for (int k = 0; k < per; ++k) push(lo[0], hi[1] - (hi[1] - lo[1]) * k / per);
  return out;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <cloud.pcd|cloud.f32> [threshold] [max_iterations] [min_plane_size] [--prob P] [--max-planes N] "
                         "[--pipeline SEED RADIUS TNUM BORDERS_OUT POST_T] [--bench STEPS]\n", argv[0]);
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
  int pos = 0, bench_steps = 0, tnum = 0;
  bool pipeline = false;
  unsigned seed = 0;
  float post_t = -1.f;
  double radius = 0;
  std::string borders_out;
  for (int i = 2; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--prob" && i + 1 < argc) det.setProbability(std::atof(argv[++i]));
    else if (a == "--max-planes" && i + 1 < argc) det.setMaxPlanes(std::atoi(argv[++i]));
    else if (a == "--bench" && i + 1 < argc) bench_steps = std::atoi(argv[++i]);
    else if (a == "--pipeline" && i + 5 < argc) {
      pipeline = true;
      seed = (unsigned)std::strtoul(argv[++i], nullptr, 10);
      radius = std::atof(argv[++i]);
      tnum = std::atoi(argv[++i]);
      borders_out = argv[++i];
      post_t = (float)std::atof(argv[++i]);
    } else if (pos == 0) { det.setDistanceThreshold(std::atof(argv[i])); ++pos; }  // config.txt T_dist_point_plane
    else if (pos == 1) { det.setMaxIterations(std::atoi(argv[i])); ++pos; }
    else if (pos == 2) { det.setMinPlaneSize(std::atoi(argv[i])); ++pos; }          // config.txt T_num_of_single_plane
  }
  const size_t n = cloud.size();

  if (bench_steps > 0) {
    PinnedCloud pinned(n);
    std::memcpy(pinned.data(), cloud.data(), n * sizeof(PointXYZ));
    std::vector<PlaneView> views;
    for (int w = 0; w < 3; ++w)
      if (!det.detectViews(pinned.data(), n, views)) { std::fprintf(stderr, "%s\n", det.last_error().c_str()); return 1; }
    double total = 0;
    for (int s = 0; s < bench_steps; ++s) {
      plane_ransac_flush_l2(det.context());
      const auto t0 = std::chrono::steady_clock::now();
      if (!det.detectViews(pinned.data(), n, views)) { std::fprintf(stderr, "%s\n", det.last_error().c_str()); return 1; }
      total += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    size_t inl = 0;
    for (const PlaneView& v : views) inl += v.size;
    std::printf("bench points %zu planes %zu inliers %zu steps %d e2e_ms_per_step %.4f\n", n, views.size(), inl, bench_steps, total / bench_steps);
    return 0;
  }

  const std::vector<PointXYZ> original = cloud;
  std::vector<PlaneRecord> planes;
  if (!det.detect(cloud, planes)) {
    std::fprintf(stderr, "%s\n", det.last_error().c_str());
    return 1;
  }
  std::printf("points %zu planes %zu remaining %zu\n", n, planes.size(), cloud.size());
  for (size_t k = 0; k < planes.size(); ++k)
    std::printf("plane %zu inliers %zu coeff %a %a %a %a hash %llu\n", k, planes[k].indices.size(), planes[k].coeff[0],
                planes[k].coeff[1], planes[k].coeff[2], planes[k].coeff[3], list_hash(planes[k].indices.data(), planes[k].indices.size()));
  if (!pipeline) return 0;

  // postProcessPlanes: re-absorption against each plane's outline, then clusterFilt, then "run again"
  std::vector<std::vector<PointXYZ>> borders;
  for (const PlaneRecord& pl : planes) borders.push_back(outline(original, pl));
  {
    std::ofstream f(borders_out, std::ios::binary);
    for (const auto& b : borders) {
      const uint32_t m = (uint32_t)b.size();
      f.write(reinterpret_cast<const char*>(&m), 4);
      f.write(reinterpret_cast<const char*>(b.data()), (std::streamsize)(b.size() * sizeof(PointXYZ)));
    }
  }
  std::vector<std::vector<int32_t>> claimed;
  if (!det.postProcessLists(planes, borders, seed, claimed, post_t)) { std::fprintf(stderr, "%s\n", det.last_error().c_str()); return 1; }
  size_t left = 0;
  if (!det.remainingCount(&left)) { std::fprintf(stderr, "%s\n", det.last_error().c_str()); return 1; }
  for (size_t k = 0; k < claimed.size(); ++k)
    std::printf("post plane %zu claimed %zu hash %llu\n", k, claimed[k].size(), list_hash(claimed[k].data(), claimed[k].size()));
  std::printf("post remaining %zu\n", left);
  size_t removed = 0;
  if (!det.clusterFilter(radius, tnum, &removed, &left)) { std::fprintf(stderr, "%s\n", det.last_error().c_str()); return 1; }
  std::printf("cluster removed %zu left %zu\n", removed, left);
  std::vector<PlaneRecord> again;
  if (!det.runAgain(again)) { std::fprintf(stderr, "%s\n", det.last_error().c_str()); return 1; }
  size_t final_left = 0;
  det.remainingCount(&final_left);
  std::printf("again planes %zu remaining %zu\n", again.size(), final_left);
  for (size_t k = 0; k < again.size(); ++k)
    std::printf("again plane %zu inliers %zu coeff %a %a %a %a hash %llu\n", k, again[k].indices.size(), again[k].coeff[0], again[k].coeff[1],
                again[k].coeff[2], again[k].coeff[3], list_hash(again[k].indices.data(), again[k].indices.size()));
  return 0;
}
