// PlaneDetectRansac.h — C++ host shim with the PlaneDetect call surface, over the C ABI in plane_ransac.h.
//
// What it stands in for in czh55/Dialog: the stages between estimateNormal() and polyPlanes() in
// PCLViewer::on_autoPerformAction_triggered (Dialog/PCLViewer.cpp:1183-1226) — createPS ... mergePlanes —
// plus postProcessPlanes' re-absorption, peel and clusterFilt (Dialog/PlaneDetect.h:1454-1656) and the
// "run again" flow (Dialog/PCLViewer.cpp:1120-1178).  Input is the cloud the reference keeps in
// `PointCloudT::Ptr source_cloud` (Dialog/PlaneDetect.h:104); output is one record per plane with the four
// coefficients the reference stores in Plane::coeff.values (Dialog/PlaneDetect.h:1493-1497) and the inlier
// indices that make up Plane::points_set (Dialog/HeaderFile.h:81-88); the cloud is replaced by the
// unclaimed points, as the reference does with source_cloud.
//
// Two levels:
//   detectViews()   the fast path: the cloud is read where it lies (a page-locked buffer — PinnedCloud below, or the
//                   caller's own storage after pinCallerMemory() — goes up in chunks that are scored as they land, see
//                   plane_ransac_set_cloud_async), the index lists come back into the shim's page-locked buffers under
//                   the later rounds, and the planes are returned as views into those buffers.  No copy of the cloud
//                   or of the lists is made on the host.
//   detect()        the same call with owning results (std::vector per plane, the cloud replaced by the remaining
//                   points), as PlaneDetect's callers expect; it copies the lists once.
// Header-only.  Define PLANE_RANSAC_WITH_PCL before including it in a PCL build to get the overloads
// on pcl::PointCloud<pcl::PointXYZ>::Ptr / pcl::ModelCoefficients / pcl::PointIndices; without it the
// same surface works on a POD cloud that is layout-compatible with pcl::PointXYZ (16 bytes).
// Errors follow the reference's convention (message + early return with an empty result,
// Dialog/PlaneDetect.h:371-375): the call returns false and last_error() holds the text.
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

// A plane of the last detectViews() call; the index arrays live in the detector's page-locked buffers and stay valid
// until its next detect call.
struct PlaneView {
  float coeff[4];
  const int32_t* indices = nullptr;      // indices into the cloud given to the call
  const int32_t* indices_cur = nullptr;  // indices into the cloud of the round
  size_t size = 0;
};

// Page-locked cloud storage (cudaMallocHost behind plane_ransac_host_alloc): what a caller fills instead of a
// std::vector when it wants the overlapped upload.
class PinnedCloud {
 public:
  PinnedCloud() = default;
  explicit PinnedCloud(size_t n) { resize(n); }
  ~PinnedCloud() { plane_ransac_host_free(p_); }
  PinnedCloud(const PinnedCloud&) = delete;
  PinnedCloud& operator=(const PinnedCloud&) = delete;
  bool resize(size_t n) {  // contents are not preserved
    if (n > cap_) {
      plane_ransac_host_free(p_);
      p_ = nullptr;
      cap_ = 0;
      void* q = nullptr;
      if (plane_ransac_host_alloc((n ? n : 1) * sizeof(PointXYZ), &q) != PR_OK) { n_ = 0; return false; }
      p_ = static_cast<PointXYZ*>(q);
      cap_ = n;
    }
    n_ = n;
    return true;
  }
  PointXYZ* data() { return p_; }
  const PointXYZ* data() const { return p_; }
  size_t size() const { return n_; }
  PointXYZ& operator[](size_t i) { return p_[i]; }

 private:
  PointXYZ* p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
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
  ~PlaneDetectRansac() {
    plane_ransac_destroy(ctx_);
    plane_ransac_host_free(cur_);
    plane_ransac_host_free(orig_);
  }
  PlaneDetectRansac(const PlaneDetectRansac&) = delete;
  PlaneDetectRansac& operator=(const PlaneDetectRansac&) = delete;

  bool ok() const { return ctx_ != nullptr; }
  const std::string& last_error() const { return err_; }
  plane_ransac_ctx* context() { return ctx_; }

  // parameters: T_dist_point_plane / T_num_of_single_plane of config.txt, and the SACSegmentation knobs
  void setDistanceThreshold(double t) { prm_.distance_threshold = t; }
  void setMaxIterations(int it) { prm_.max_iterations = it; }
  void setMinPlaneSize(int n) { prm_.min_plane_size = n; }
  void setProbability(double p) { prm_.probability = p; }
  void setOptimizeCoefficients(bool on) { prm_.optimize_coefficients = on ? 1 : 0; }
  void setMaxPlanes(int n) { prm_.max_planes = n; }
  void setDotOrder(int order) { prm_.dot_order = order; }
  void setScorer(int scorer) { prm_.scorer = scorer; }
  void setRefitMode(int mode) { prm_.refit_mode = mode; }  // PR_REFIT_FIXED (default) / PR_REFIT_PCL_FLOAT
  const pr_params& params() const { return prm_; }

  // Page-locks storage the caller owns (e.g. pcl::PointCloud::points) in place, so that detect calls on it take the
  // overlapped upload; unpin before the storage is freed or reallocated.  Pinning costs milliseconds per 100 MB: worth
  // it for a cloud that is detected on more than once or that a reader fills right after.
  bool pinCallerMemory(void* p, size_t bytes) { return plane_ransac_host_register(p, bytes) == PR_OK || fail(); }
  bool unpinCallerMemory(void* p) { return plane_ransac_host_unregister(p) == PR_OK || fail(); }

  // ---- the fast path ---------------------------------------------------------------------------------------------
  // cloud: n points, read in place (page-locked: overlapped upload; pageable: one blocking copy); not modified.
  // planes: views into this object's page-locked index buffers.  The unclaimed points stay on the device: fetch them
  // with remainingCount() / remaining(), or go on with postProcessViews / clusterFilter / runAgain.
  bool detectViews(const PointXYZ* cloud, size_t n, std::vector<PlaneView>& planes) {
    planes.clear();
    if (!ctx_) return false;
    if (plane_ransac_set_cloud_async(ctx_, reinterpret_cast<const pr_point*>(cloud), n) != PR_OK) return fail();
    const int mp = prm_.max_planes > 0 ? prm_.max_planes : 0;
    coeffs_.assign(4 * (size_t)(mp ? mp : 1), 0.f);
    offs_.assign((size_t)mp + 1, 0);
    if (!reserveLists(n)) return false;
    int found = 0;
    if (plane_ransac_extract_planes(ctx_, &prm_, coeffs_.data(), cur_, orig_, n, offs_.data(), &found, nullptr) != PR_OK)
      return fail();
    planes.resize((size_t)found);
    for (int k = 0; k < found; ++k) {
      std::memcpy(planes[k].coeff, &coeffs_[4 * (size_t)k], 4 * sizeof(float));
      planes[k].indices = orig_ + offs_[k];
      planes[k].indices_cur = cur_ + offs_[k];
      planes[k].size = offs_[k + 1] - offs_[k];
    }
    return true;
  }

  // Points no plane (and no later pass) has claimed, in their original order.
  bool remainingCount(size_t* n) {
    if (!ctx_) return false;
    return plane_ransac_remaining(ctx_, nullptr, 0, n) == PR_OK || fail();
  }
  bool remaining(PointXYZ* out, size_t cap, size_t* n) {
    if (!ctx_) return false;
    return plane_ransac_remaining(ctx_, reinterpret_cast<pr_point*>(out), cap, n) == PR_OK || fail();
  }

  // ---- owning results, as PlaneDetect's callers expect -----------------------------------------------------------------
  // cloud in, planes out; `cloud` is replaced by the points no plane claimed (original order).
  bool detect(std::vector<PointXYZ>& cloud, std::vector<PlaneRecord>& planes) {
    planes.clear();
    std::vector<PlaneView> views;
    if (!detectViews(cloud.data(), cloud.size(), views)) return false;
    planes.resize(views.size());
    for (size_t k = 0; k < views.size(); ++k) {
      std::memcpy(planes[k].coeff, views[k].coeff, sizeof(views[k].coeff));
      planes[k].indices.assign(views[k].indices, views[k].indices + views[k].size);
      planes[k].indices_cur.assign(views[k].indices_cur, views[k].indices_cur + views[k].size);
    }
    return fetchRemaining(cloud);
  }

  // postProcessPlanes' second half (Dialog/PlaneDetect.h:1530-1566) after a detect call: every point the device still
  // holds as unclaimed is tested against the planes' polygons (Plane::border, one per record in `borders`) with the
  // reference's isPointInPoly; a claimed point is appended to every plane that contains it (indices into the cloud
  // given to detect()) and `cloud` receives the points no polygon claimed (whatever it held before is discarded: the
  // pass works on the device's copy).  rand_seed: the reference's srand(time(0)).
  bool postProcess(std::vector<PointXYZ>& cloud, std::vector<PlaneRecord>& planes,
                   const std::vector<std::vector<PointXYZ>>& borders, unsigned rand_seed, float dist_threshold = -1.f) {
    std::vector<std::vector<int32_t>> claimed;
    if (!postProcessLists(planes, borders, rand_seed, claimed, dist_threshold)) return false;
    for (size_t k = 0; k < planes.size(); ++k) planes[k].indices.insert(planes[k].indices.end(), claimed[k].begin(), claimed[k].end());
    return fetchRemaining(cloud);
  }

  // The same pass without fetching the leftovers (they stay on the device for clusterFilter / runAgain / remaining()):
  // claimed[k] = indices (into the cloud given to detect) of the points plane k's polygon claimed.  dist_threshold:
  // isPointInPoly's T_dist_point_plane (non-strict); negative = the detection threshold, as in the reference where both
  // read the same config.txt entry.
  bool postProcessLists(const std::vector<PlaneRecord>& planes, const std::vector<std::vector<PointXYZ>>& borders,
                        unsigned rand_seed, std::vector<std::vector<int32_t>>& claimed, float dist_threshold = -1.f) {
    if (!ctx_) return false;
    const float T = dist_threshold < 0.f ? (float)prm_.distance_threshold : dist_threshold;
    const size_t P = planes.size();
    if (borders.size() != P) { err_ = "one border polygon per plane"; return false; }
    claimed.assign(P, std::vector<int32_t>());
    std::vector<float> coeffs(4 * (P ? P : 1));
    std::vector<PointXYZ> bd;
    std::vector<size_t> bo(P + 1, 0), po(P + 1, 0);
    for (size_t k = 0; k < P; ++k) {
      std::memcpy(&coeffs[4 * k], planes[k].coeff, 4 * sizeof(float));
      bd.insert(bd.end(), borders[k].begin(), borders[k].end());
      bo[k + 1] = bd.size();
    }
    // sized from what the device holds, not from any vector of the caller: usually no more points are claimed than
    // are left; a point inside several polygons is listed once per plane, hence the retry with the worst case
    size_t n_staged = 0, n_cur = 0, n_rem = 0;
    if (plane_ransac_cloud_size(ctx_, &n_staged, &n_cur) != PR_OK) return fail();
    size_t cap = n_cur ? n_cur : 1;
    std::vector<int32_t> orig(cap);
    int rc = plane_ransac_reabsorb(ctx_, coeffs.data(), reinterpret_cast<const pr_point*>(bd.data()), bo.data(), (int)P,
                                   T, rand_seed, nullptr, orig.data(), cap, po.data(), &n_rem);
    if (rc == PR_ERR_CAPACITY) {  // nothing was modified: run again with room for every (point, plane) pair
      cap *= (P ? P : 1);
      orig.resize(cap);
      rc = plane_ransac_reabsorb(ctx_, coeffs.data(), reinterpret_cast<const pr_point*>(bd.data()), bo.data(), (int)P,
                                 T, rand_seed, nullptr, orig.data(), cap, po.data(), &n_rem);
    }
    if (rc != PR_OK) return fail();
    for (size_t k = 0; k < P; ++k) claimed[k].assign(orig.begin() + po[k], orig.begin() + po[k + 1]);
    return true;
  }

  // clusterFilt() (Dialog/PlaneDetect.h:1582-1656), the last step of postProcessPlanes, on the device's leftovers:
  // connected components of the radius graph with at most max_small_cluster points (T_cluster_num) are dropped.
  bool clusterFilter(double radius, int max_small_cluster, size_t* n_removed, size_t* n_left) {
    if (!ctx_) return false;
    return plane_ransac_cluster_filter(ctx_, radius, max_small_cluster, n_removed, n_left) == PR_OK || fail();
  }

  // "Run again" (PCLViewer::on_runAgainAction_triggered, Dialog/PCLViewer.cpp:1120-1178): detection on what the
  // previous passes left, without a round trip through the host.  Indices in `planes` refer to the cloud given to the
  // first detect call (the library composes the index maps).
  bool runAgain(std::vector<PlaneRecord>& planes) {
    planes.clear();
    if (!ctx_) return false;
    if (plane_ransac_restage_remaining(ctx_) != PR_OK) return fail();
    size_t n = 0;
    if (plane_ransac_cloud_size(ctx_, &n, nullptr) != PR_OK) return fail();
    const int mp = prm_.max_planes > 0 ? prm_.max_planes : 0;
    coeffs_.assign(4 * (size_t)(mp ? mp : 1), 0.f);
    offs_.assign((size_t)mp + 1, 0);
    if (!reserveLists(n)) return false;
    int found = 0;
    if (plane_ransac_extract_planes(ctx_, &prm_, coeffs_.data(), cur_, orig_, n, offs_.data(), &found, nullptr) != PR_OK)
      return fail();
    std::vector<int32_t> map(n ? n : 1);  // restaged point -> index in the first call's cloud
    if (plane_ransac_staged_source_indices(ctx_, map.data(), n) != PR_OK) return fail();
    planes.resize((size_t)found);
    for (int k = 0; k < found; ++k) {
      std::memcpy(planes[k].coeff, &coeffs_[4 * (size_t)k], 4 * sizeof(float));
      planes[k].indices_cur.assign(cur_ + offs_[k], cur_ + offs_[k + 1]);
      planes[k].indices.resize(offs_[k + 1] - offs_[k]);
      for (size_t i = offs_[k]; i < offs_[k + 1]; ++i) planes[k].indices[i - offs_[k]] = map[(size_t)orig_[i]];
    }
    return true;
  }

  // The unclaimed points into `cloud` (replacing its contents).
  bool fetchRemaining(std::vector<PointXYZ>& cloud) {
    size_t n_rem = 0;
    if (!remainingCount(&n_rem)) return false;
    std::vector<PointXYZ> rest(n_rem);
    if (n_rem && !remaining(rest.data(), n_rem, &n_rem)) return false;
    cloud.swap(rest);
    return true;
  }

#ifdef PLANE_RANSAC_WITH_PCL
  // The north-star surface: (PointCloud<PointXYZ>::Ptr, threshold, max iterations, min plane size) ->
  // coefficients + inlier indices per plane; *cloud is replaced by the remaining points.  The points are read where
  // they lie (pinCallerMemory(cloud->points.data(), ...) beforehand gives the overlapped upload) and the leftovers are
  // written straight into cloud->points: no intermediate copy of the cloud.
  bool detect(pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud, std::vector<pcl::ModelCoefficients>& coefficients,
              std::vector<pcl::PointIndices>& inliers) {
    static_assert(sizeof(pcl::PointXYZ) == sizeof(pr_point), "pcl::PointXYZ layout");
    coefficients.clear();
    inliers.clear();
    std::vector<PlaneView> views;
    if (!detectViews(reinterpret_cast<const PointXYZ*>(cloud->points.data()), cloud->points.size(), views)) return false;
    coefficients.resize(views.size());
    inliers.resize(views.size());
    for (size_t k = 0; k < views.size(); ++k) {
      coefficients[k].values.assign(views[k].coeff, views[k].coeff + 4);
      inliers[k].indices.assign(views[k].indices, views[k].indices + views[k].size);
    }
    size_t n_rem = 0;
    if (!remainingCount(&n_rem)) return false;
    cloud->points.resize(n_rem);  // shrinks in place: the storage (and a pinning of it) stays
    if (n_rem && !remaining(reinterpret_cast<PointXYZ*>(cloud->points.data()), n_rem, &n_rem)) return false;
    cloud->width = (uint32_t)n_rem;
    cloud->height = 1;
    return true;
  }
#endif

 private:
  bool fail() {
    err_ = plane_ransac_last_error();
    return false;
  }
  // page-locked index buffers, kept across calls: each plane's lists land there while the later rounds still score
  bool reserveLists(size_t n) {
    if (n <= list_cap_ && cur_) return true;
    plane_ransac_host_free(cur_);
    plane_ransac_host_free(orig_);
    cur_ = orig_ = nullptr;
    list_cap_ = 0;
    void *a = nullptr, *b = nullptr;
    const size_t bytes = (n ? n : 1) * sizeof(int32_t);
    if (plane_ransac_host_alloc(bytes, &a) != PR_OK || plane_ransac_host_alloc(bytes, &b) != PR_OK) {
      plane_ransac_host_free(a);
      return fail();
    }
    cur_ = static_cast<int32_t*>(a);
    orig_ = static_cast<int32_t*>(b);
    list_cap_ = n;
    return true;
  }
  plane_ransac_ctx* ctx_ = nullptr;
  pr_params prm_;
  std::string err_;
  int32_t* cur_ = nullptr;
  int32_t* orig_ = nullptr;
  size_t list_cap_ = 0;
  std::vector<float> coeffs_;
  std::vector<size_t> offs_;
};

}  // namespace plane_detect_ransac
