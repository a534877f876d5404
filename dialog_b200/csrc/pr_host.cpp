// pr_host.cpp — see pr_host.hpp.  Restates PCL 1.8 host-side behaviour (sac_model.h, ransac.hpp,
// common/impl/eigen.hpp); PCL itself is not part of the reference tree.
#include "pr_host.hpp"

#include "pr_draw.h"
#include "pr_math.h"

#include <cfloat>
#include <climits>
#include <cmath>
#include <algorithm>
#include <cstring>

namespace pr {

void Mt19937::seed_with(uint32_t seed) {
  mt_[0] = seed;
  for (int i = 1; i < 624; ++i) mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + (uint32_t)i;
  idx_ = 624;
}

uint32_t Mt19937::next() {
  if (idx_ >= 624) {
    auto twist = [this](int i, int i1, int im) {
      const uint32_t y = (mt_[i] & 0x80000000u) | (mt_[i1] & 0x7fffffffu);
      mt_[i] = mt_[im] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    };
    for (int i = 0; i < 227; ++i) twist(i, i + 1, i + 397);
    for (int i = 227; i < 623; ++i) twist(i, i + 1, i - 227);
    twist(623, 0, 396);
    idx_ = 0;
  }
  uint32_t y = mt_[idx_++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

static constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;

void IndexSampler::extend_raw(size_t upto) {
  while (raw_.size() < upto) {
    const size_t at = raw_.size();
    raw_.resize(at + 1248);
    for (size_t k = at; k < raw_.size(); ++k) raw_[k] = rng_.next() >> 1;
  }
}

void IndexSampler::reset(size_t n, uint32_t seed) {
  if (!raw_valid_ || raw_seed_ != seed) {
    rng_.seed_with(seed);
    raw_.clear();
    raw_seed_ = seed;
    raw_valid_ = true;
  }
  pos_ = 0;
  n_ = n;
  head_[0] = 0;
  head_[1] = 1;
  head_[2] = 2;
  if (slots_.empty()) {
    slots_.assign(1u << 12, Slot{kEmptyKey, 0});
    mask_ = (1u << 12) - 1;
  } else if (touched_.size() * 4 > slots_.size()) {
    std::fill(slots_.begin(), slots_.end(), Slot{kEmptyKey, 0});
  } else {
    for (uint32_t h : touched_) slots_[h].key = kEmptyKey;
  }
  touched_.clear();
  for (int i = 0; i < 3; ++i) {
    const uint64_t d = n > (size_t)i ? (uint64_t)(n - i) : 1;
    d_[i] = d <= 0xFFFFFFFFull ? (uint32_t)d : 0u;  // 0: divisor does not fit 32 bits, use the plain operator
    m_[i] = d_[i] ? UINT64_C(0xFFFFFFFFFFFFFFFF) / d_[i] + 1 : 0;
  }
}

void IndexSampler::reserve(size_t draws) {
  if (!touched_.empty()) return;
  uint32_t cap = mask_ + 1;
  while ((size_t)cap < draws * 3 * 4 && cap < (1u << 28)) cap <<= 1;  // load factor <= 1/4
  if (cap == mask_ + 1) return;
  slots_.assign(cap, Slot{kEmptyKey, 0});
  mask_ = cap - 1;
}

static inline uint32_t hash_index(uint32_t k) { return (k * 2654435761u) >> 7; }

void IndexSampler::grow() {
  std::vector<Slot> old;
  old.swap(slots_);
  const uint32_t cap = (mask_ + 1) * 4;
  slots_.assign(cap, Slot{kEmptyKey, 0});
  mask_ = cap - 1;
  touched_.clear();
  for (const Slot& s : old)
    if (s.key != kEmptyKey) exchange(s.key, s.val);
}

int32_t IndexSampler::exchange(uint32_t j, int32_t v) {
  for (uint32_t h = hash_index(j) & mask_;; h = (h + 1) & mask_) {
    Slot& s = slots_[h];
    if (s.key == j) {
      const int32_t old = s.val;
      s.val = v;
      return old;
    }
    if (s.key == kEmptyKey) {
      s.key = j;
      s.val = v;
      touched_.push_back(h);
      if (touched_.size() * 3 > mask_) grow();
      return (int32_t)j;  // untouched entries hold the identity
    }
  }
}

inline size_t IndexSampler::index_of(uint32_t i, uint32_t r) const {
  if (d_[i]) {
    const uint64_t low = m_[i] * r;  // r % d_[i]
    return i + (size_t)(((unsigned __int128)low * d_[i]) >> 64);
  }
  return i + (size_t)r % (n_ - i);
}

void IndexSampler::draw(int32_t out[3]) {
  if (pos_ + 6 > raw_.size()) extend_raw(pos_ + 6);
  const uint32_t* rr = raw_.data() + pos_;
  pos_ += 3;
  // the table slots of the NEXT draw: its indices depend only on the random stream, not on the permutation
  for (uint32_t i = 0; i < 3; ++i) {
    const size_t jn = index_of(i, rr[3 + i]);
    if (jn >= 3) __builtin_prefetch(&slots_[hash_index((uint32_t)jn) & mask_], 1, 1);
  }
  for (uint32_t i = 0; i < 3; ++i) {
    // rnd() = boost::uniform_int<>(0, INT_MAX) over mt19937 == rng() >> 1
    const size_t j = index_of(i, rr[i]);
    if (j < 3) {
      const int32_t t = head_[i];
      head_[i] = head_[j];
      head_[j] = t;
    } else {
      head_[i] = exchange((uint32_t)j, head_[i]);
    }
  }
  out[0] = head_[0];
  out[1] = head_[1];
  out[2] = head_[2];
}

void fill_rnd_stream(uint32_t seed, size_t count, uint32_t* out) {
  Mt19937 rng(seed);
  for (size_t k = 0; k < count; ++k) out[k] = rng.next() >> 1;
}

namespace {
struct HostAtomics {
  static unsigned long long cas(unsigned long long* p, unsigned long long expect, unsigned long long v) {
    const unsigned long long old = *p;
    if (old == expect) *p = v;
    return old;
  }
  static unsigned long long load(const unsigned long long* p) { return *p; }
  static uint32_t add(uint32_t* p, uint32_t v) {
    const uint32_t old = *p;
    *p += v;
    return old;
  }
};
}  // namespace

bool draw_triples_parallel(size_t n_points, uint32_t seed, int n_draws, int32_t* triples) {
  const size_t n_ops = 3 * (size_t)n_draws;
  if (n_ops == 0) return true;
  std::vector<uint32_t> rnd(n_ops);
  fill_rnd_stream(seed, n_ops, rnd.data());
  size_t cap = 1024;
  while (cap < 4 * n_ops) cap <<= 1;
  std::vector<unsigned long long> table(cap, 0ull);  // epoch 0 = never used
  std::vector<uint32_t> coll(kDrawCollCap);
  uint32_t n_coll = 0;
  // the device runs the ops in any order: emulate the reverse one, so that the op that owns a table slot is never
  // the earliest op of its position
  for (size_t k = n_ops; k-- > 0;)
    draw_scatter<HostAtomics>((uint32_t)k, rnd[k], (uint32_t)n_points, triples, table.data(), (uint32_t)(cap - 1), 1u, coll.data(), &n_coll,
                              (uint32_t)coll.size());
  if (n_coll > (uint32_t)coll.size()) return false;
  std::sort(coll.begin(), coll.begin() + n_coll);
  const size_t distinct = (size_t)(std::unique(coll.begin(), coll.begin() + n_coll) - coll.begin());
  if (distinct > (size_t)kDrawMaxCollisions) return false;
  uint32_t map_cap = 16;
  while (map_cap < 2 * distinct) map_cap <<= 1;
  std::vector<uint32_t> mk(map_cap, kDrawNoOp);
  std::vector<int32_t> mv(map_cap, 0);
  // the kernel's two paths: independent groups resolved one entry per thread when that is exact, else the sequential replay
  const std::vector<int32_t> before(triples, triples + n_ops);  // v as the parallel phase left it (the kernel prefetches it)
  auto fetch = [&before](int, uint32_t idx) { return before[idx]; };
  bool independent = true;
  for (size_t i = 0; i < distinct; ++i) independent = independent && draw_independent_ok(coll.data(), (int)i, fetch);
  if (independent) {
    for (size_t i = 0; i < distinct; ++i) triples[coll[i]] = draw_resolve_independent(coll.data(), (int)i, fetch);
    return true;
  }
  draw_resolve(coll.data(), (int)distinct, triples, [triples](int, uint32_t idx) { return triples[idx]; }, mk.data(), mv.data(), map_cap - 1);
  return true;
}

RansacReplay::RansacReplay(long long n_points, int max_iterations, double probability)
    : one_over_n_(1.0 / (double)n_points),
      log_probability_(std::log(1.0 - probability)),
      max_iterations_(max_iterations),
      n_best_(-INT_MAX),
      max_skip_((unsigned)max_iterations * 10u) {
  done_ = !loop_condition();
}

bool RansacReplay::loop_condition() const { return (double)iterations_ < k_ && skipped_ < max_skip_; }

int RansacReplay::draws_wanted() const {
  if (done_) return 0;
  // at most (max_iterations + 1) trials are ever scored; k may cut that short
  long long left = (long long)max_iterations_ + 1 - iterations_;
  if (k_ < (double)left + iterations_) {
    double kk = std::ceil(k_) - iterations_;
    if (kk < (double)left) left = (long long)kk;
  }
  if (left < 1) left = 1;
  return (int)left;
}

bool RansacReplay::feed(const int32_t* counts, const uint8_t* good, int n) {
  for (int j = 0; j < n && !done_; ++j) {
    ++draws_used_;
    if (!good[j]) {
      // getSamples: redraw; after max_sample_checks_ (1000) failures the selection is empty and
      // computeModel breaks out of its loop.
      if (++bad_run_ >= 1000) done_ = true;
      continue;
    }
    bad_run_ = 0;
    // computeModelCoefficients cannot fail on a sample isSampleGood accepted (same test), so
    // skipped_count stays 0 for the plane model; kept for fidelity with the loop condition.
    const int c = counts[j];
    if (c > n_best_) {
      n_best_ = c;
      best_draw_ = draws_used_ - 1;
      const double w = (double)n_best_ * one_over_n_;
      double p_no_outliers = 1.0 - std::pow(w, 3.0);
      if (p_no_outliers < DBL_EPSILON) p_no_outliers = DBL_EPSILON;
      if (p_no_outliers > 1.0 - DBL_EPSILON) p_no_outliers = 1.0 - DBL_EPSILON;
      k_ = log_probability_ / std::log(p_no_outliers);
    }
    ++iterations_;
    if (iterations_ > max_iterations_) done_ = true;
    if (!loop_condition()) done_ = true;
  }
  return done_;
}

float key_to_float(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
  float f;
  std::memcpy(&f, &b, 4);
  return f;
}

uint32_t float_to_key(float f) {
  uint32_t b;
  std::memcpy(&b, &f, 4);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

bool centroid_from_sums(const long long sums[4], const float lo[3], int scale_exp, float centroid[3]) {
  if (sums[3] <= 0) return false;
  const double inv = std::ldexp(1.0, -scale_exp);
  for (int a = 0; a < 3; ++a)
    centroid[a] = (float)((double)lo[a] + ((double)sums[a] / (double)sums[3]) * inv);
  return true;
}

int scale_exp_from_bbox_keys(const uint32_t keys[6]) {
  double r = 0.0;
  for (int a = 0; a < 3; ++a) {
    if (keys[a] > keys[3 + a]) continue;  // no finite point
    const double e = (double)key_to_float(keys[3 + a]) - (double)key_to_float(keys[a]);
    if (e > r) r = e;
  }
  if (!(r > 0.0)) return 0;
  int e;
  (void)std::frexp(r, &e);
  return 30 - e;
}

// --- pcl::eigen33 / computeRoots / computeRoots2 (common/impl/eigen.hpp) in double: pr_math.h, the same source the
// device-side round loop runs (portable atan2 / sin / cos, so host and kernel agree bit for bit) ---------------------
bool plane_from_moments(const int64_t m[16], const float pivot[3], int scale_exp, float coeff[4]) {
  long long mm[16];
  for (int i = 0; i < 16; ++i) mm[i] = (long long)m[i];
  return pm_plane_from_moments(mm, pivot, scale_exp, coeff);
}

// --- pcl::eigen33 / computeRoots / computeRoots2 in FP32, as PCL 1.8 instantiates them for the plane model ------------
namespace {
void roots2f(float b, float c, float roots[3]) {
  roots[0] = 0.f;
  float d = (float)(b * b - 4.0 * c);  // PCL writes the literal 4.0: the product is formed in double
  if (d < 0.0) d = 0.f;
  const float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

void roots3f(const float m[9], float roots[3]) {
  const float c0 = m[0] * m[4] * m[8] + 2.f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
  const float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  const float c2 = m[0] + m[4] + m[8];
  if (std::fabs(c0) < FLT_EPSILON) {
    roots2f(c2, c1, roots);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = std::sqrt(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f) a_over_3 = 0.f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f) q = 0.f;
  const float rho = std::sqrt(-a_over_3);
  const float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
  const float cos_theta = std::cos(theta);
  const float sin_theta = std::sin(theta);
  roots[0] = c2_over_3 + 2.f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  float tmp;
  if (roots[0] >= roots[1]) { tmp = roots[0]; roots[0] = roots[1]; roots[1] = tmp; }
  if (roots[1] >= roots[2]) {
    tmp = roots[1]; roots[1] = roots[2]; roots[2] = tmp;
    if (roots[0] >= roots[1]) { tmp = roots[0]; roots[0] = roots[1]; roots[1] = tmp; }
  }
  if (roots[0] <= 0) roots2f(c2, c1, roots);
}
}  // namespace

bool plane_from_pcl_float_sums(const float sums[9], long long n_inliers, float coeff[4]) {
  if (n_inliers < 4) return false;  // PCL: "if (inliers.size () <= 3)" -> the input coefficients stand
  float accu[9];
  const float cnt = (float)n_inliers;
  for (int i = 0; i < 9; ++i) accu[i] = sums[i] / cnt;
  float s[9];
  s[0] = accu[0] - accu[6] * accu[6];
  s[1] = accu[1] - accu[6] * accu[7];
  s[2] = accu[2] - accu[6] * accu[8];
  s[4] = accu[3] - accu[7] * accu[7];
  s[5] = accu[4] - accu[7] * accu[8];
  s[8] = accu[5] - accu[8] * accu[8];
  s[3] = s[1];
  s[6] = s[2];
  s[7] = s[5];
  float scale = 0.f;
  for (int i = 0; i < 9; ++i) {
    const float a = std::fabs(s[i]);
    if (a > scale) scale = a;
  }
  if (scale <= FLT_MIN) scale = 1.f;
  for (int i = 0; i < 9; ++i) s[i] = s[i] / scale;
  float ev[3];
  roots3f(s, ev);
  s[0] -= ev[0];
  s[4] -= ev[0];
  s[8] -= ev[0];
  const float v1[3] = {s[1] * s[5] - s[2] * s[4], s[2] * s[3] - s[0] * s[5], s[0] * s[4] - s[1] * s[3]};
  const float v2[3] = {s[1] * s[8] - s[2] * s[7], s[2] * s[6] - s[0] * s[8], s[0] * s[7] - s[1] * s[6]};
  const float v3[3] = {s[4] * s[8] - s[5] * s[7], s[5] * s[6] - s[3] * s[8], s[3] * s[7] - s[4] * s[6]};
  const float len1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  const float len2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  const float len3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* best;
  float len;
  if (len1 >= len2 && len1 >= len3) { best = v1; len = len1; }
  else if (len2 >= len1 && len2 >= len3) { best = v2; len = len2; }
  else { best = v3; len = len3; }
  const float nrm = std::sqrt(len);
  const float v[3] = {best[0] / nrm, best[1] / nrm, best[2] / nrm};
  // Hessian form: d = -(v . centroid), the 4-term dot in Eigen's SSE2 order with a zero fourth coefficient
  const float dot = (v[0] * accu[6] + v[2] * accu[8]) + (v[1] * accu[7] + 0.0f * 1.0f);
  coeff[0] = v[0];
  coeff[1] = v[1];
  coeff[2] = v[2];
  coeff[3] = -1.0f * dot;
  return true;
}

float threshold_up(double t) {
  float f = (float)t;  // round to nearest
  if ((double)f < t) f = std::nextafterf(f, INFINITY);
  return f;
}

void msvc_rand_edges(unsigned seed, int border_size, int32_t edges[10]) {
  uint32_t hold = seed;
  for (int i = 0; i < 10; ++i) {
    hold = hold * 214013u + 2531011u;
    const uint32_t r = (hold >> 16) & 0x7fffu;
    edges[i] = (int32_t)(r % (uint32_t)border_size);
  }
}

void shard_range(long long n_points, int n_ranks, int rank, long long* first, long long* count) {
  const long long base = n_points / n_ranks, rem = n_points % n_ranks;
  *first = base * rank + (rank < rem ? rank : rem);
  *count = base + (rank < rem ? 1 : 0);
}

}  // namespace pr
