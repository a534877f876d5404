// pr_host.hpp — host-side logic of the plane-RANSAC backend: PCL's sampling stream, the sequential
// RANSAC decision loop replayed over batched device counts, and the closed-form plane from integer
// moments.  Pure C++ (no CUDA), compiled with -ffp-contract=off.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace pr {

// boost::mt19937 / std::mt19937 (PCL sac_model.h: rng_alg_).
class Mt19937 {
 public:
  explicit Mt19937(uint32_t seed = 5489u) { seed_with(seed); }
  void seed_with(uint32_t seed);
  uint32_t next();

 private:
  uint32_t mt_[624];
  int idx_ = 624;
};

// SampleConsensusModel::drawIndexSample for a 3-point model (PCL 1.8 sac_model.h).  The permutation
// shuffled_indices_ starts as the identity over n indices and persists across draws; only the
// entries touched so far are stored, so a 10^8-point cloud costs nothing to (re)initialise.
class IndexSampler {
 public:
  IndexSampler() : n_(0) {}
  IndexSampler(size_t n, uint32_t seed) { reset(n, seed); }
  // Start over for a cloud of n points (fresh RNG, identity permutation); the table allocation is kept and only the
  // slots touched since the last reset are cleared, so a peel round costs microseconds whatever the table size.
  void reset(size_t n, uint32_t seed);
  void draw(int32_t out[3]);
  // Size the table for this many draws up front (each draw touches up to three entries).
  void reserve(size_t draws);
  size_t size() const { return n_; }

 private:
  // shuffled_indices_[j] for j >= 3: open-addressing table (key and value side by side: one cache line per probe)
  // of the entries that differ from the identity
  struct Slot { uint32_t key; int32_t val; };
  int32_t exchange(uint32_t j, int32_t v);  // old = shuffled[j]; shuffled[j] = v; return old
  void grow();
  void extend_raw(size_t upto);
  inline size_t index_of(uint32_t i, uint32_t r) const;
  // rnd() values (rng() >> 1).  PCL reseeds the generator for every segment() call, so every peel round consumes
  // the same stream: it is generated once per seed and replayed; knowing the next values also lets draw() prefetch
  // the table slots of the following draw.
  Mt19937 rng_;
  std::vector<uint32_t> raw_;
  uint32_t raw_seed_ = 0;
  bool raw_valid_ = false;
  size_t pos_ = 0;
  size_t n_;
  int32_t head_[3];  // shuffled_indices_[0..3), touched by every swap
  std::vector<Slot> slots_;
  std::vector<uint32_t> touched_;
  uint32_t mask_ = 0;
  // r % (n - i) by multiplication (Lemire's fastmod, exact for 32-bit operands): m_[i] = 2^64 / (n - i) + 1
  uint64_t m_[3] = {0, 0, 0};
  uint32_t d_[3] = {1, 1, 1};
};

// The first `count` values of rnd() = mt19937(seed)() >> 1, the stream drawIndexSample consumes.
void fill_rnd_stream(uint32_t seed, size_t count, uint32_t* out);
// The parallel formulation of the sampler the device-side round loop runs (pr_draw.h), emulated on the host with the
// same per-op code: triples of the first n_draws draws over n_points indices.  Returns false when more than
// kDrawMaxCollisions ops would have to be replayed sequentially (the device then leaves the round to the host sampler).
bool draw_triples_parallel(size_t n_points, uint32_t seed, int n_draws, int32_t* triples);

// RandomSampleConsensus::computeModel's while-loop (PCL 1.8 ransac.hpp), fed one draw at a time.
// The draw stream does not depend on the scores, so the device scores a batch of draws and this
// object replays PCL's sequential decisions (strict '>' first-best, adaptive k, iteration cap,
// getSamples' 1000-redraw limit) over them.
class RansacReplay {
 public:
  RansacReplay(long long n_points, int max_iterations, double probability);
  // Consumes draws [0, n); returns true when the loop has terminated (no more draws needed).
  bool feed(const int32_t* counts, const uint8_t* good, int n);
  bool done() const { return done_; }
  // Upper bound on the draws the loop can still consume if every one of them is good.
  int draws_wanted() const;
  int best_draw() const { return best_draw_; }  // global draw index, -1 if none
  int best_count() const { return n_best_; }
  int iterations() const { return iterations_; }
  int draws_used() const { return draws_used_; }
  int skipped() const { return (int)skipped_; }

 private:
  bool loop_condition() const;
  double one_over_n_, log_probability_, k_ = 1.0;
  int max_iterations_, iterations_ = 0, n_best_, best_draw_ = -1, draws_used_ = 0, bad_run_ = 0;
  unsigned skipped_ = 0, max_skip_;
  bool done_ = false;
};

// Extent exponent s: |x - pivot| * 2^s < 2^30 for all finite points of a cloud with the given
// bounding box (ordered-float keys as written by the staging kernel: min x,y,z then max x,y,z).
int scale_exp_from_bbox_keys(const uint32_t keys[6]);

// Ordered-float key <-> float (keys compare like the floats they encode).
uint32_t float_to_key(float f);
float key_to_float(uint32_t k);
// Exact centroid of the finite points from integer coordinate sums about `lo` on the 2^-scale_exp grid
// (sums[3] = number of points); false when there are none.
bool centroid_from_sums(const long long sums[4], const float lo[3], int scale_exp, float centroid[3]);

// Least-squares plane from exact integer moments (DESIGN.md "refit").  Returns false (coeff
// untouched) when fewer than 4 points contributed.
bool plane_from_moments(const int64_t m[16], const float pivot[3], int scale_exp, float coeff[4]);

// PCL 1.8's float path of optimizeModelCoefficients after the nine sequential sums (PR_REFIT_PCL_FLOAT):
// computeMeanAndCovarianceMatrix's division and covariance, pcl::eigen33 in FP32 (sqrtf / atan2f / cosf / sinf of the
// platform, as PCL itself calls them), Hessian d.  accu: xx, xy, xz, yy, yz, zz, x, y, z.  Fewer than 4 inliers:
// coeff keeps the sample's model (false).
bool plane_from_pcl_float_sums(const float accu[9], long long n_inliers, float coeff[4]);

// Smallest float >= t: the FP32 threshold equivalent to PCL's float-vs-double strict compare.
float threshold_up(double t);

void shard_range(long long n_points, int n_ranks, int rank, long long* first, long long* count);

// isPointInPoly's edge draws (Dialog/PlaneDetect.h:1905-1919): srand(seed), then ten times rand() % border_size with
// the MSVC CRT generator (holdrand = holdrand * 214013 + 2531011; (holdrand >> 16) & 0x7fff).
void msvc_rand_edges(unsigned seed, int border_size, int32_t edges[10]);

}  // namespace pr
