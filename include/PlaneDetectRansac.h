// PlaneDetectRansac.h — C++ host shim with the PlaneDetect call surface, over the C ABI in plane_ransac.h.
//
// What it stands in for in czh55/Dialog: the stages between estimateNormal() and polyPlanes() in
// PCLViewer::on_autoPerformAction_triggered (Dialog/PCLViewer.cpp:1183-1226) — createPS ... mergePlanes —
// plus the peel in postProcessPlanes (Dialog/PlaneDetect.h:1560-1566).  Input is the cloud the reference
// keeps in `PointCloudT::Ptr source_cloud` (Dialog/PlaneDetect.h:104); output is one record per plane
// with the four coefficients the reference stores in Plane::coeff.values (Dialog/PlaneDetect.h:1493-1497)
// and the inlier indices that make up Plane::points_set (Dialog/HeaderFile.h:81-88); the cloud is replaced
// by the unclaimed points, as the reference does with source_cloud.
//
// Header-only.  Define PLANE_RANSAC_WITH_PCL before including it in a PCL build to get the overloads
// on pcl::PointCloud<pcl::PointXYZ>::Ptr / pcl::ModelCoefficients / pcl::PointIndices; without it the
// same surface works on a POD cloud that is layout-compatible with pcl::PointXYZ (16 bytes).
// Errors follow the reference's convention (message + early return with an empty result,
// Dialog/PlaneDetect.h:371-375): detect() returns false and last_error() holds the text.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "plane_ransac.h"

#ifdef PLANE_RANSAC_WITH_PCL
#include <pcl/ModelCoefficients.h>
#include <pcl/PointIndices.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#endif

namespace plane_detect_ransac {

struct PointXYZ {  // == pcl::PointXYZ: x, y, z + 4 bytes of padding
  float x, y, z, pad;
};
static_assert(sizeof(PointXYZ) == sizeof(pr_point), "PointXYZ must be 16 bytes like pcl::PointXYZ");

struct PlaneRecord {
  float coeff[4];                    // a, b, c, d  (== Plane::coeff.values)
  std::vector<int32_t> indices;      // inliers as indices into the cloud given to detect()
  std::vector<int32_t> indices_cur;  // the same points as indices into the cloud of the round (PCL's loop)
};

class PlaneDetectRansac {
 public:
  explicit PlaneDetectRansac(int device = 0) {
    if (plane_ransac_create(&ctx_, device) != PR_OK) {
      err_ = plane_ransac_last_error();
      ctx_ = nullptr;
    }
    plane_ransac_default_params(&prm_);
  }
  ~PlaneDetectRansac() { plane_ransac_destroy(ctx_); }
  PlaneDetectRansac(const PlaneDetectRansac&) = delete;
  PlaneDetectRansac& operator=(const PlaneDetectRansac&) = delete;

  bool ok() const { return ctx_ != nullptr; }
  const std::string& last_error() const { return err_; }

  // parameters: T_dist_point_plane / T_num_of_single_plane of config.txt, and the SACSegmentation knobs
  void setDistanceThreshold(double t) { prm_.distance_threshold = t; }
  void setMaxIterations(int it) { prm_.max_iterations = it; }
  void setMinPlaneSize(int n) { prm_.min_plane_size = n; }
  void setProbability(double p) { prm_.probability = p; }
  void setOptimizeCoefficients(bool on) { prm_.optimize_coefficients = on ? 1 : 0; }
  void setMaxPlanes(int n) { prm_.max_planes = n; }
  void setDotOrder(int order) { prm_.dot_order = order; }
  void setScorer(int scorer) { prm_.scorer = scorer; }
  const pr_params& params() const { return prm_; }

  // cloud in, planes out; `cloud` is replaced by the points no plane claimed (original order).
  bool detect(std::vector<PointXYZ>& cloud, std::vector<PlaneRecord>& planes) {
    planes.clear();
    if (!ctx_) return false;
    const size_t n = cloud.size();
    if (plane_ransac_set_cloud(ctx_, reinterpret_cast<const pr_point*>(cloud.data()), n) != PR_OK) return fail();
    const int mp = prm_.max_planes > 0 ? prm_.max_planes : 0;
    std::vector<float> coeffs(4 * (size_t)(mp ? mp : 1));
    std::vector<int32_t> cur(n ? n : 1), orig(n ? n : 1);
    std::vector<size_t> offs((size_t)mp + 1, 0);
    int found = 0;
    if (plane_ransac_extract_planes(ctx_, &prm_, coeffs.data(), cur.data(), orig.data(), n, offs.data(), &found,
                                    nullptr) != PR_OK)
      return fail();
    planes.resize((size_t)found);
    for (int k = 0; k < found; ++k) {
      for (int i = 0; i < 4; ++i) planes[k].coeff[i] = coeffs[4 * k + i];
      planes[k].indices.assign(orig.begin() + offs[k], orig.begin() + offs[k + 1]);
      planes[k].indices_cur.assign(cur.begin() + offs[k], cur.begin() + offs[k + 1]);
    }
    size_t n_rem = 0;
    if (plane_ransac_remaining(ctx_, nullptr, 0, &n_rem) != PR_OK) return fail();
    std::vector<PointXYZ> rest(n_rem);
    if (n_rem && plane_ransac_remaining(ctx_, reinterpret_cast<pr_point*>(rest.data()), n_rem, &n_rem) != PR_OK)
      return fail();
    cloud.swap(rest);
    return true;
  }

  // postProcessPlanes' second half (Dialog/PlaneDetect.h:1530-1566) after detect(): every point still in `cloud` is
  // tested against the planes' polygons (Plane::border, one per record in `borders`) with the reference's
  // isPointInPoly; a claimed point is appended to every plane that contains it (indices into the cloud given to
  // detect()) and `cloud` is replaced by the points no polygon claimed.  rand_seed: the reference's srand(time(0)).
  bool postProcess(std::vector<PointXYZ>& cloud, std::vector<PlaneRecord>& planes,
                   const std::vector<std::vector<PointXYZ>>& borders, unsigned rand_seed) {
    if (!ctx_) return false;
    const size_t P = planes.size();
    if (borders.size() != P) { err_ = "one border polygon per plane"; return false; }
    std::vector<float> coeffs(4 * (P ? P : 1));
    std::vector<PointXYZ> bd;
    std::vector<size_t> bo(P + 1, 0), po(P + 1, 0);
    for (size_t k = 0; k < P; ++k) {
      std::memcpy(&coeffs[4 * k], planes[k].coeff, 4 * sizeof(float));
      bd.insert(bd.end(), borders[k].begin(), borders[k].end());
      bo[k + 1] = bd.size();
    }
    size_t cap = cloud.size() ? cloud.size() : 1, n_rem = 0;
    std::vector<int32_t> orig(cap);
    int rc = plane_ransac_reabsorb(ctx_, coeffs.data(), reinterpret_cast<const pr_point*>(bd.data()), bo.data(), (int)P,
                                   (float)prm_.distance_threshold, rand_seed, nullptr, orig.data(), cap, po.data(), &n_rem);
    if (rc == PR_ERR_CAPACITY) {  // points claimed by several planes: the lists can outgrow the cloud
      cap *= (P ? P : 1);
      orig.resize(cap);
      rc = plane_ransac_reabsorb(ctx_, coeffs.data(), reinterpret_cast<const pr_point*>(bd.data()), bo.data(), (int)P,
                                 (float)prm_.distance_threshold, rand_seed, nullptr, orig.data(), cap, po.data(), &n_rem);
    }
    if (rc != PR_OK) return fail();
    for (size_t k = 0; k < P; ++k) planes[k].indices.insert(planes[k].indices.end(), orig.begin() + po[k], orig.begin() + po[k + 1]);
    std::vector<PointXYZ> rest(n_rem);
    if (n_rem && plane_ransac_remaining(ctx_, reinterpret_cast<pr_point*>(rest.data()), n_rem, &n_rem) != PR_OK) return fail();
    cloud.swap(rest);
    return true;
  }

#ifdef PLANE_RANSAC_WITH_PCL
  // The north-star surface: (PointCloud<PointXYZ>::Ptr, threshold, max iterations, min plane size) ->
  // coefficients + inlier indices per plane; *cloud is replaced by the remaining points.
  bool detect(pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud, std::vector<pcl::ModelCoefficients>& coefficients,
              std::vector<pcl::PointIndices>& inliers) {
    static_assert(sizeof(pcl::PointXYZ) == sizeof(pr_point), "pcl::PointXYZ layout");
    std::vector<PointXYZ> pts(cloud->points.size());
    std::memcpy(pts.data(), cloud->points.data(), pts.size() * sizeof(PointXYZ));
    std::vector<PlaneRecord> planes;
    if (!detect(pts, planes)) return false;
    coefficients.resize(planes.size());
    inliers.resize(planes.size());
    for (size_t k = 0; k < planes.size(); ++k) {
      coefficients[k].values.assign(planes[k].coeff, planes[k].coeff + 4);
      inliers[k].indices.assign(planes[k].indices.begin(), planes[k].indices.end());
    }
    cloud->points.resize(pts.size());
    std::memcpy(cloud->points.data(), pts.data(), pts.size() * sizeof(PointXYZ));
    cloud->width = (uint32_t)pts.size();
    cloud->height = 1;
    return true;
  }
#endif

 private:
  bool fail() {
    err_ = plane_ransac_last_error();
    return false;
  }
  plane_ransac_ctx* ctx_ = nullptr;
  pr_params prm_;
  std::string err_;
};

}  // namespace plane_detect_ransac
