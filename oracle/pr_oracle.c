/*
 * pr_oracle.c — CPU oracle for the plane-RANSAC hot path.  TEST INFRASTRUCTURE ONLY (see pr_oracle.h).
 *
 * PARITY UNPINNED against the reference: the algorithm lives in PCL 1.8 (un-vendored dependency of
 * czh55/Dialog, Dialog/PropertySheet-success.props:6,11), which is absent here; the reference has no
 * tests or golden vectors for it.  Each function restates the published PCL 1.8 behaviour of the
 * source file it names (restated from the PCL 1.8 sources, not copied from /root/reference, which
 * does not contain them).  Known places where PCL's result depends on its build and this file had
 * to pick one behaviour are marked "CHOICE".  All of them, in one place:
 *   - rnd() == rng() >> 1: boost::uniform_int<>(0, INT_MAX) over mt19937.  Holds for Boost >= 1.47 (the
 *     generate_uniform_int algorithm: engine range 2^32 - 1 over target range 2^31 - 1 gives bucket_size 2 and never
 *     rejects); Boost is not vendored with the reference, PCL 1.8.x binaries for MSVC v140 shipped with Boost 1.6x.
 *   - Eigen >= 3.3: Vector::normalize() and "accu /= n" DIVIDE (Eigen 3.2 multiplied by the reciprocal).
 *   - Eigen's 4-wide FP32 reductions run in SSE2 order (e0 + e2) + (e1 + e3) (MSVC v140 x64, no AVX): the plane dot
 *     product of countWithinDistance (ORC_DOT_PCL_SSE2), the squared norm and d of computeModelCoefficients, the
 *     Hessian d of optimizeModelCoefficients.  ORC_DOT_FMA is this backend's own faster order, also implemented.
 *   - no FP contraction on the reference toolchain (MSVC v140 /fp:precise): every '*' and '+' rounds separately.
 *   - the line model's optimised axis (SACMODEL_LINE, only used to replay the reference's own SAC call) comes from a
 *     double-precision Jacobi instead of PCL's FP32 eigen33 + computeCorrespondingEigenVector.
 *   - the libm behind eigen33's atan2 / cos / sin is this platform's (glibc), not MSVC's CRT.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -fPIC -shared  (see oracle/Makefile).
 * -ffp-contract=off keeps every '*' and '+' separately rounded (MSVC v140 / SSE2 code generation);
 * fused multiply-adds appear only where fmaf() is written.
 */
#include "pr_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* =========================================================================================
 * boost::mt19937 (bit-identical to std::mt19937).  PCL: sac_model.h, rng_alg_.seed(12345u).
 * ========================================================================================= */
void orc_mt_seed(orc_mt19937* g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

uint32_t orc_mt_next(orc_mt19937* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
      if (y & 1u) v ^= 0x9908b0dfu;
      g->mt[i] = v;
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* =========================================================================================
 * SampleConsensusModel::rnd / drawIndexSample (sac_model.h).
 * rnd() is boost::uniform_int<>(0, INT_MAX) over mt19937: engine range 2^32-1, target range
 * 2^31-1 => bucket_size 2, never rejects => rnd() == rng() >> 1.
 * shuffled_indices_ persists across draws; a fresh model (fresh seed, fresh identity permutation)
 * is built by every segment() call.
 * ========================================================================================= */
int orc_sampler_init(orc_sampler* s, size_t n, uint32_t seed) {
  orc_mt_seed(&s->rng, seed);
  s->n = n;
  s->shuffled = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
  if (!s->shuffled) return -1;
  for (size_t i = 0; i < n; ++i) s->shuffled[i] = (int32_t)i;
  return 0;
}

void orc_sampler_free(orc_sampler* s) {
  free(s->shuffled);
  s->shuffled = NULL;
}

/* drawIndexSample for a model of sample_size points (plane: 3, line: 2) */
void orc_sampler_draw_k(orc_sampler* s, int sample_size, int32_t* idx) {
  for (size_t i = 0; i < (size_t)sample_size; ++i) {
    uint32_t r = orc_mt_next(&s->rng) >> 1; /* rnd() */
    size_t j = i + (size_t)r % (s->n - i);
    int32_t tmp = s->shuffled[i];
    s->shuffled[i] = s->shuffled[j];
    s->shuffled[j] = tmp;
  }
  for (int i = 0; i < sample_size; ++i) idx[i] = s->shuffled[i];
}

void orc_sampler_draw(orc_sampler* s, int32_t idx[3]) { orc_sampler_draw_k(s, 3, idx); }

int orc_draw_sequence(size_t n, uint32_t seed, int n_draws, int32_t* triples) {
  if (n < 3) return -1;
  orc_sampler s;
  if (orc_sampler_init(&s, n, seed)) return -1;
  for (int k = 0; k < n_draws; ++k) orc_sampler_draw(&s, triples + 3 * k);
  orc_sampler_free(&s);
  return 0;
}

/* =========================================================================================
 * SampleConsensusModelPlane (sac_model_plane.hpp)
 * ========================================================================================= */

/* isSampleGood: dy1dy2 = (p1-p0)/(p2-p0) on the Array4f maps (x,y,z,pad); good iff
 * dy1dy2[0] != dy1dy2[1] || dy1dy2[2] != dy1dy2[1].  NaN (0/0) compares unequal => good. */
int orc_is_sample_good(const orc_point* cloud, const int32_t idx[3]) {
  const orc_point *p0 = &cloud[idx[0]], *p1 = &cloud[idx[1]], *p2 = &cloud[idx[2]];
  float r0 = (p1->x - p0->x) / (p2->x - p0->x);
  float r1 = (p1->y - p0->y) / (p2->y - p0->y);
  float r2 = (p1->z - p0->z) / (p2->z - p0->z);
  return (r0 != r1) || (r2 != r1);
}

/* computeModelCoefficients: collinearity test, cross product, Eigen normalize(), d = -(n·p0).
 * CHOICE: Eigen normalize() divides by the norm (Eigen >= 3.3 semantics; 3.2 multiplied by 1/norm).
 * squaredNorm() and the 4-term dot for d use the SSE2 reduction order (e0 + e2) + (e1 + e3). */
int orc_compute_model(const orc_point* cloud, const int32_t idx[3], float coeff[4]) {
  const orc_point *p0 = &cloud[idx[0]], *p1 = &cloud[idx[1]], *p2 = &cloud[idx[2]];
  float ux = p1->x - p0->x, uy = p1->y - p0->y, uz = p1->z - p0->z;
  float vx = p2->x - p0->x, vy = p2->y - p0->y, vz = p2->z - p0->z;
  float r0 = ux / vx, r1 = uy / vy, r2 = uz / vz;
  if ((r0 == r1) && (r2 == r1)) return 0;
  float nx = uy * vz - uz * vy;
  float ny = uz * vx - ux * vz;
  float nz = ux * vy - uy * vx;
  float sq = (nx * nx + nz * nz) + (ny * ny + 0.0f * 0.0f);
  float norm = sqrtf(sq);
  nx = nx / norm;
  ny = ny / norm;
  nz = nz / norm;
  /* model_coefficients[3] = -1 * (model_coefficients.template head<4>().dot(p0.matrix())), with
   * model_coefficients[3] == 0 and p0.w == the point's padding (1.0f in pcl::PointXYZ). */
  float dot = (nx * p0->x + nz * p0->z) + (ny * p0->y + 0.0f * 1.0f);
  coeff[0] = nx;
  coeff[1] = ny;
  coeff[2] = nz;
  coeff[3] = -1.0f * dot;
  return 1;
}

static inline float dot_pcl_sse2(const float c[4], float x, float y, float z) {
  /* Eigen Vector4f::dot on SSE2: pmul then predux = (p0 + p2) + (p1 + p3), p3 = c[3] * 1.0f. */
  return (c[0] * x + c[2] * z) + (c[1] * y + c[3]);
}

static inline float dot_fma(const float c[4], float x, float y, float z) {
  return fmaf(c[0], x, fmaf(c[1], y, fmaf(c[2], z, c[3])));
}

float orc_signed_distance(const float coeff[4], const orc_point* p, int dot_order) {
  return dot_order == ORC_DOT_FMA ? dot_fma(coeff, p->x, p->y, p->z) : dot_pcl_sse2(coeff, p->x, p->y, p->z);
}

void orc_residuals(const orc_point* cloud, size_t n, const float coeff[4], int dot_order, float* out) {
  for (size_t i = 0; i < n; ++i) out[i] = orc_signed_distance(coeff, &cloud[i], dot_order);
}

/* countWithinDistance: if (fabs(coeff.dot(pt)) < threshold) ++n — float dot promoted to double for a
 * strict '<' against the double threshold. */
int64_t orc_count_within(const orc_point* cloud, size_t n, const float coeff[4], double t, int dot_order) {
  int64_t cnt = 0;
  if (dot_order == ORC_DOT_FMA) {
    for (size_t i = 0; i < n; ++i)
      if (fabs((double)dot_fma(coeff, cloud[i].x, cloud[i].y, cloud[i].z)) < t) ++cnt;
  } else {
    for (size_t i = 0; i < n; ++i)
      if (fabs((double)dot_pcl_sse2(coeff, cloud[i].x, cloud[i].y, cloud[i].z)) < t) ++cnt;
  }
  return cnt;
}

int64_t orc_count_within_mt(const orc_point* cloud, size_t n, const float coeff[4], double t, int dot_order) {
  int64_t cnt = 0;
  long long nn = (long long)n;
  if (dot_order == ORC_DOT_FMA) {
#pragma omp parallel for reduction(+ : cnt) schedule(static)
    for (long long i = 0; i < nn; ++i)
      if (fabs((double)dot_fma(coeff, cloud[i].x, cloud[i].y, cloud[i].z)) < t) ++cnt;
  } else {
#pragma omp parallel for reduction(+ : cnt) schedule(static)
    for (long long i = 0; i < nn; ++i)
      if (fabs((double)dot_pcl_sse2(coeff, cloud[i].x, cloud[i].y, cloud[i].z)) < t) ++cnt;
  }
  return cnt;
}

void orc_count_batch(const orc_point* cloud, size_t n, const float* coeffs, int K, double t, int dot_order,
                     int threads, int32_t* counts) {
  for (int k = 0; k < K; ++k)
    counts[k] = (int32_t)(threads > 1 ? orc_count_within_mt(cloud, n, coeffs + 4 * k, t, dot_order)
                                      : orc_count_within(cloud, n, coeffs + 4 * k, t, dot_order));
}

/* selectWithinDistance: same test, inliers in ascending index order. */
size_t orc_select_within(const orc_point* cloud, size_t n, const float coeff[4], double t, int dot_order,
                         int32_t* out) {
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    float r = orc_signed_distance(coeff, &cloud[i], dot_order);
    if (fabs((double)r) < t) out[m++] = (int32_t)i;
  }
  return m;
}

/* =========================================================================================
 * eigen33 / computeRoots / computeRoots2 (common/impl/eigen.hpp), float (PCL) and double.
 * ========================================================================================= */
#define ORC_EIGEN_IMPL(SUFFIX, T, EPS, TMIN, SQRT, ATAN2, COS, SIN, FABS)                                 \
  static void roots2_##SUFFIX(T b, T c, T roots[3]) {                                                   \
    roots[0] = (T)0;                                                                                     \
    T d = (T)(b * b - 4.0 * c);                                                                          \
    if (d < 0.0) d = (T)0.0;                                                                             \
    T sd = SQRT(d);                                                                                      \
    roots[2] = (T)0.5 * (b + sd);                                                                        \
    roots[1] = (T)0.5 * (b - sd);                                                                        \
  }                                                                                                      \
  static void roots_##SUFFIX(const T m[9], T roots[3]) {                                                \
    /* characteristic equation x^3 - c2 x^2 + c1 x - c0 = 0 of the symmetric matrix m (row-major) */     \
    T c0 = m[0] * m[4] * m[8] + (T)2 * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] -    \
           m[8] * m[1] * m[1];                                                                           \
    T c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];            \
    T c2 = m[0] + m[4] + m[8];                                                                           \
    if (FABS(c0) < EPS) {                                                                                \
      roots2_##SUFFIX(c2, c1, roots);                                                                    \
      return;                                                                                            \
    }                                                                                                    \
    const T s_inv3 = (T)(1.0 / 3.0);                                                                     \
    const T s_sqrt3 = SQRT((T)3.0);                                                                      \
    T c2_over_3 = c2 * s_inv3;                                                                           \
    T a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;                                                         \
    if (a_over_3 > (T)0) a_over_3 = (T)0;                                                                \
    T half_b = (T)0.5 * (c0 + c2_over_3 * ((T)2 * c2_over_3 * c2_over_3 - c1));                          \
    T q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;                                              \
    if (q > (T)0) q = (T)0;                                                                              \
    T rho = SQRT(-a_over_3);                                                                             \
    T theta = ATAN2(SQRT(-q), half_b) * s_inv3;                                                          \
    T cos_theta = COS(theta);                                                                            \
    T sin_theta = SIN(theta);                                                                            \
    roots[0] = c2_over_3 + (T)2 * rho * cos_theta;                                                       \
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);                                      \
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);                                      \
    T tmp;                                                                                               \
    if (roots[0] >= roots[1]) { tmp = roots[0]; roots[0] = roots[1]; roots[1] = tmp; }                   \
    if (roots[1] >= roots[2]) {                                                                          \
      tmp = roots[1]; roots[1] = roots[2]; roots[2] = tmp;                                               \
      if (roots[0] >= roots[1]) { tmp = roots[0]; roots[0] = roots[1]; roots[1] = tmp; }                 \
    }                                                                                                    \
    if (roots[0] <= 0) roots2_##SUFFIX(c2, c1, roots);                                                   \
  }                                                                                                      \
  /* eigen33(mat, eigenvalue, eigenvector): eigenvector of the smallest eigenvalue */                   \
  static void eigen33_##SUFFIX(const T mat[9], T* eigenvalue, T vec[3]) {                               \
    T scale = (T)0;                                                                                      \
    for (int i = 0; i < 9; ++i) { T a = FABS(mat[i]); if (a > scale) scale = a; }                        \
    if (scale <= TMIN) scale = (T)1.0;                                                                   \
    T s[9];                                                                                              \
    for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;                                                   \
    T ev[3];                                                                                             \
    roots_##SUFFIX(s, ev);                                                                               \
    *eigenvalue = ev[0] * scale;                                                                         \
    s[0] -= ev[0]; s[4] -= ev[0]; s[8] -= ev[0];                                                         \
    const T *r0 = s, *r1 = s + 3, *r2 = s + 6;                                                           \
    T v1[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]}; \
    T v2[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]}; \
    T v3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]}; \
    T len1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];                                              \
    T len2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];                                              \
    T len3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];                                              \
    const T* best; T len;                                                                                \
    if (len1 >= len2 && len1 >= len3) { best = v1; len = len1; }                                         \
    else if (len2 >= len1 && len2 >= len3) { best = v2; len = len2; }                                    \
    else { best = v3; len = len3; }                                                                      \
    T nrm = SQRT(len);                                                                                   \
    vec[0] = best[0] / nrm; vec[1] = best[1] / nrm; vec[2] = best[2] / nrm;                              \
  }

ORC_EIGEN_IMPL(f, float, FLT_EPSILON, FLT_MIN, sqrtf, atan2f, cosf, sinf, fabsf)
ORC_EIGEN_IMPL(d, double, DBL_EPSILON, DBL_MIN, sqrt, atan2, cos, sin, fabs)

/* =========================================================================================
 * optimizeModelCoefficients, PCL 1.8 float path: computeMeanAndCovarianceMatrix (centroid.hpp,
 * indices overload, dense cloud) + eigen33 + Hessian d.
 * CHOICE: "accu /= point_count" divides (Eigen >= 3.3); the 4-term dot for d is in SSE2 order.
 * ========================================================================================= */
int orc_refit_pcl_float(const orc_point* cloud, const int32_t* idx, size_t n_idx, const float coeff_in[4],
                        float coeff_out[4]) {
  if (n_idx < 4) { /* PCL: "if (inliers.size () <= 3)" -> optimized = input */
    memcpy(coeff_out, coeff_in, 4 * sizeof(float));
    return 0;
  }
  float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t k = 0; k < n_idx; ++k) {
    const orc_point* p = &cloud[idx[k]];
    accu[0] += p->x * p->x;
    accu[1] += p->x * p->y;
    accu[2] += p->x * p->z;
    accu[3] += p->y * p->y;
    accu[4] += p->y * p->z;
    accu[5] += p->z * p->z;
    accu[6] += p->x;
    accu[7] += p->y;
    accu[8] += p->z;
  }
  float cnt = (float)n_idx;
  for (int i = 0; i < 9; ++i) accu[i] = accu[i] / cnt;
  float cov[9];
  cov[0] = accu[0] - accu[6] * accu[6];
  cov[1] = accu[1] - accu[6] * accu[7];
  cov[2] = accu[2] - accu[6] * accu[8];
  cov[4] = accu[3] - accu[7] * accu[7];
  cov[5] = accu[4] - accu[7] * accu[8];
  cov[8] = accu[5] - accu[8] * accu[8];
  cov[3] = cov[1];
  cov[6] = cov[2];
  cov[7] = cov[5];
  float ev, v[3];
  eigen33_f(cov, &ev, v);
  float dot = (v[0] * accu[6] + v[2] * accu[8]) + (v[1] * accu[7] + 0.0f * 1.0f);
  float out[4] = {v[0], v[1], v[2], -1.0f * dot};
  /* isModelValid(): the plane model only checks the coefficient count; a NaN eigenvector passes
   * it in PCL too, so it is passed through here. */
  memcpy(coeff_out, out, sizeof(out));
  return 1;
}

/* =========================================================================================
 * optimizeModelCoefficients, order-independent form (ORC_REFIT_FIXED).
 *
 *   s       : extent exponent of the staged cloud, |x - pivot| * 2^s < 2^30 for all finite points
 *   q_a     : llrint(((double)p.a - (double)pivot.a) * 2^s)            (exact up to the final rint)
 *   moments : n, S_a = sum q_a, S_ab = sum q_a q_b                      (exact integers)
 *   C_ab    : n * S_ab - S_a * S_b                                      (exact, 128-bit)
 *   plane   : eigen33 in double on (double)C_ab; centroid = pivot + (S_a / n) * 2^-s;
 *             d = -((nx*cx + ny*cy) + nz*cz); coefficients rounded to float.
 *
 * Integer sums commute, so any traversal order, thread count or GPU count gives the same bits.
 * The quantisation step is 2^-30 of the cloud extent, 64x finer than one float ulp at that extent.
 * ========================================================================================= */
int orc_fixed_scale_exp(const orc_point* cloud, size_t n) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (size_t i = 0; i < n; ++i) {
    const float c[3] = {cloud[i].x, cloud[i].y, cloud[i].z};
    if (!(isfinite(c[0]) && isfinite(c[1]) && isfinite(c[2]))) continue;
    for (int a = 0; a < 3; ++a) {
      if (c[a] < lo[a]) lo[a] = c[a];
      if (c[a] > hi[a]) hi[a] = c[a];
    }
  }
  double r = 0.0;
  for (int a = 0; a < 3; ++a) {
    if (hi[a] >= lo[a]) {
      double e = (double)hi[a] - (double)lo[a];
      if (e > r) r = e;
    }
  }
  if (!(r > 0.0)) return 0;
  int e;
  (void)frexp(r, &e); /* r = m * 2^e, m in [0.5, 1)  =>  r < 2^e */
  return 30 - e;
}

size_t orc_remove_nonfinite(const orc_point* cloud, size_t n, orc_point* out, int32_t* map) {
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!(isfinite(cloud[i].x) && isfinite(cloud[i].y) && isfinite(cloud[i].z))) continue;
    if (map) map[m] = (int32_t)i;
    out[m++] = cloud[i];
  }
  return m;
}

int orc_centroid_exact(const orc_point* cloud, size_t n, float centroid[3]) {
  float lo[3] = {INFINITY, INFINITY, INFINITY};
  size_t cnt = 0;
  for (size_t i = 0; i < n; ++i) {
    const float c[3] = {cloud[i].x, cloud[i].y, cloud[i].z};
    if (!(isfinite(c[0]) && isfinite(c[1]) && isfinite(c[2]))) continue;
    ++cnt;
    for (int a = 0; a < 3; ++a)
      if (c[a] < lo[a]) lo[a] = c[a];
  }
  centroid[0] = centroid[1] = centroid[2] = 0.0f;
  if (!cnt) return 0;
  const int s = orc_fixed_scale_exp(cloud, n);
  const double sc = ldexp(1.0, s), inv = ldexp(1.0, -s);
  long long sum[3] = {0, 0, 0};
  for (size_t i = 0; i < n; ++i) {
    const float c[3] = {cloud[i].x, cloud[i].y, cloud[i].z};
    if (!(isfinite(c[0]) && isfinite(c[1]) && isfinite(c[2]))) continue;
    for (int a = 0; a < 3; ++a) sum[a] += llrint(((double)c[a] - (double)lo[a]) * sc);
  }
  for (int a = 0; a < 3; ++a) centroid[a] = (float)((double)lo[a] + ((double)sum[a] / (double)cnt) * inv);
  return 1;
}

void orc_translate(orc_point* cloud, size_t n, const float centroid[3]) {
  for (size_t i = 0; i < n; ++i) {
    cloud[i].x -= centroid[0];
    cloud[i].y -= centroid[1];
    cloud[i].z -= centroid[2];
  }
}

void orc_project_points(const orc_point* cloud, const int32_t* idx, size_t n_idx, const float coeff[4], orc_point* out) {
  for (size_t i = 0; i < n_idx; ++i) {
    const orc_point* p = &cloud[idx[i]];
    float lambda = 2.0 * (coeff[0] * p->x + coeff[1] * p->y + coeff[2] * p->z + coeff[3]);
    out[i].x = p->x - lambda / 2.0 * coeff[0];
    out[i].y = p->y - lambda / 2.0 * coeff[1];
    out[i].z = p->z - lambda / 2.0 * coeff[2];
    out[i].w = 1.0f;
  }
}

typedef __int128 i128;

int orc_plane_from_moments(const int64_t m[16], const float pivot[3], int scale_exp, float coeff_out[4]) {
  int64_t n = m[0];
  if (n < 4) return 0;
  i128 S[3] = {m[1], m[2], m[3]};
  i128 Sab[6];
  for (int k = 0; k < 6; ++k) Sab[k] = (i128)m[4 + 2 * k] * ((i128)1 << 32) + (i128)m[5 + 2 * k];
  /* order: xx, xy, xz, yy, yz, zz */
  static const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {0, 1, 2, 1, 2, 2};
  double C[6];
  for (int k = 0; k < 6; ++k) C[k] = (double)((i128)n * Sab[k] - S[A[k]] * S[B[k]]);
  double cov[9] = {C[0], C[1], C[2], C[1], C[3], C[4], C[2], C[4], C[5]};
  double ev, v[3];
  eigen33_d(cov, &ev, v);
  double inv = ldexp(1.0, -scale_exp);
  double cx = (double)pivot[0] + ((double)m[1] / (double)n) * inv;
  double cy = (double)pivot[1] + ((double)m[2] / (double)n) * inv;
  double cz = (double)pivot[2] + ((double)m[3] / (double)n) * inv;
  double d = -((v[0] * cx + v[1] * cy) + v[2] * cz);
  coeff_out[0] = (float)v[0];
  coeff_out[1] = (float)v[1];
  coeff_out[2] = (float)v[2];
  coeff_out[3] = (float)d;
  return 1;
}

int orc_refit_fixed(const orc_point* cloud, const int32_t* idx, size_t n_idx, const float pivot[3],
                    int scale_exp, const float coeff_in[4], float coeff_out[4], int64_t* moments_out) {
  int64_t m[16];
  memset(m, 0, sizeof(m));
  if (n_idx >= 4) {
    double sc = ldexp(1.0, scale_exp);
    i128 S[3] = {0, 0, 0}, Sab[6] = {0, 0, 0, 0, 0, 0};
    for (size_t k = 0; k < n_idx; ++k) {
      const orc_point* p = &cloud[idx[k]];
      long long qx = llrint(((double)p->x - (double)pivot[0]) * sc);
      long long qy = llrint(((double)p->y - (double)pivot[1]) * sc);
      long long qz = llrint(((double)p->z - (double)pivot[2]) * sc);
      S[0] += qx; S[1] += qy; S[2] += qz;
      Sab[0] += (i128)qx * qx; Sab[1] += (i128)qx * qy; Sab[2] += (i128)qx * qz;
      Sab[3] += (i128)qy * qy; Sab[4] += (i128)qy * qz; Sab[5] += (i128)qz * qz;
    }
    m[0] = (int64_t)n_idx;
    m[1] = (int64_t)S[0]; m[2] = (int64_t)S[1]; m[3] = (int64_t)S[2];
    for (int k = 0; k < 6; ++k) {
      /* S_ab = hi * 2^32 + lo with 0 <= lo < 2^32 (floor split; same value as any other split) */
      i128 hi = Sab[k] >> 32;
      m[4 + 2 * k] = (int64_t)hi;
      m[5 + 2 * k] = (int64_t)(Sab[k] - hi * ((i128)1 << 32));
    }
  }
  if (moments_out) memcpy(moments_out, m, sizeof(m));
  if (n_idx < 4 || !orc_plane_from_moments(m, pivot, scale_exp, coeff_out)) {
    memcpy(coeff_out, coeff_in, 4 * sizeof(float));
    return 0;
  }
  return 1;
}

/* =========================================================================================
 * RandomSampleConsensus::computeModel (sample_consensus/impl/ransac.hpp) followed by
 * SACSegmentation::segment (segmentation/impl/sac_segmentation.hpp).
 * ========================================================================================= */
/* =========================================================================================
 * RandomSampleConsensus::computeModel (sample_consensus/impl/ransac.hpp) over a model given by its sample size and its
 * three callbacks — the SAME loop (and the same sampler) serves the plane model of the hot path and the line model of
 * the reference's only literal SAC call (Dialog/SimplifyVerticesSize.cpp:64-67).
 * ========================================================================================= */
typedef struct {
  int sample_size;
  int (*is_sample_good)(const orc_point* cloud, const int32_t* idx);
  int (*compute_model)(const orc_point* cloud, const int32_t* idx, float* coeff);
  int64_t (*count_within)(const orc_point* cloud, size_t n, const float* coeff, double t, int dot_order);
  int n_coeff;
} orc_sac_model;

typedef struct {
  int iterations, draws, skipped, have_model, n_best;
  int32_t best_sample[3];
  float best_coeff[6];
} orc_sac_result;

static int orc_ransac_compute_model(const orc_point* cloud, size_t n, const orc_params* prm, const orc_sac_model* mdl, orc_sac_result* out) {
  memset(out, 0, sizeof(*out));
  const double threshold = prm->distance_threshold;
  const int max_iterations = prm->max_iterations;
  int iterations = 0;
  int n_best = -INT_MAX;
  double k = 1.0;
  const double log_probability = log(1.0 - prm->probability);
  const double one_over_indices = 1.0 / (double)n;
  unsigned skipped = 0;
  const unsigned max_skip = (unsigned)max_iterations * 10u;
  const unsigned max_sample_checks = 1000;
  const size_t ss = (size_t)mdl->sample_size;

  orc_sampler smp;
  smp.shuffled = NULL;
  if (n >= ss && orc_sampler_init(&smp, n, prm->seed)) return -1;

  while ((double)iterations < k && skipped < max_skip) {
    /* getSamples: fewer points than the sample size => empty selection, loop ends */
    int32_t sel[3] = {0, 0, 0};
    int got = 0;
    if (n >= ss) {
      for (unsigned it = 0; it < max_sample_checks; ++it) {
        orc_sampler_draw_k(&smp, mdl->sample_size, sel);
        ++out->draws;
        if (mdl->is_sample_good(cloud, sel)) { got = 1; break; }
      }
    }
    if (!got) break; /* "No samples could be selected!" */
    float mc[6];
    if (!mdl->compute_model(cloud, sel, mc)) {
      ++skipped;
      continue;
    }
    int cnt = (int)mdl->count_within(cloud, n, mc, threshold, prm->dot_order);
    if (cnt > n_best) {
      n_best = cnt;
      memcpy(out->best_sample, sel, sizeof(sel));
      memcpy(out->best_coeff, mc, (size_t)mdl->n_coeff * sizeof(float));
      out->have_model = 1;
      double w = (double)n_best * one_over_indices;
      double p_no_outliers = 1.0 - pow(w, (double)mdl->sample_size);
      if (p_no_outliers < DBL_EPSILON) p_no_outliers = DBL_EPSILON;             /* (std::max)(eps, p) */
      if (p_no_outliers > 1.0 - DBL_EPSILON) p_no_outliers = 1.0 - DBL_EPSILON; /* (std::min)(1-eps, p) */
      k = log_probability / log(p_no_outliers);
    }
    ++iterations;
    if (iterations > max_iterations) break;
  }
  if (n >= ss) orc_sampler_free(&smp);
  out->iterations = iterations;
  out->skipped = (int)skipped;
  out->n_best = n_best;
  return 0;
}

static int plane_good_cb(const orc_point* cloud, const int32_t* idx) { return orc_is_sample_good(cloud, idx); }
static int plane_model_cb(const orc_point* cloud, const int32_t* idx, float* coeff) { return orc_compute_model(cloud, idx, coeff); }
static int64_t plane_count_cb(const orc_point* cloud, size_t n, const float* coeff, double t, int dot_order) {
  return orc_count_within(cloud, n, coeff, t, dot_order);
}

int orc_segment(const orc_point* cloud, size_t n, const orc_params* prm, int scale_exp_or_min,
                float coeff[4], int32_t* inliers, size_t* n_inliers, orc_trace* trace) {
  orc_trace tr;
  memset(&tr, 0, sizeof(tr));
  *n_inliers = 0;
  coeff[0] = coeff[1] = coeff[2] = coeff[3] = 0.0f;
  const double threshold = prm->distance_threshold;
  const orc_sac_model plane = {3, plane_good_cb, plane_model_cb, plane_count_cb, 4};
  orc_sac_result res;
  if (orc_ransac_compute_model(cloud, n, prm, &plane, &res)) return -1;
  const int iterations = res.iterations, have_model = res.have_model, n_best = res.n_best;
  const unsigned skipped = (unsigned)res.skipped;
  tr.draws = res.draws;
  int32_t best_sample[3];
  float best_coeff[4];
  memcpy(best_sample, res.best_sample, sizeof(best_sample));
  memcpy(best_coeff, res.best_coeff, sizeof(best_coeff));

  tr.iterations = iterations;
  tr.skipped = (int)skipped;
  tr.ok = have_model;
  tr.scale_exp = scale_exp_or_min;
  if (!have_model) { /* segment(): "Error segmenting the model! No solution found." -> outputs cleared */
    if (trace) *trace = tr;
    return 0;
  }
  memcpy(tr.best_sample, best_sample, sizeof(best_sample));
  memcpy(tr.raw_coeff, best_coeff, sizeof(best_coeff));
  tr.best_count = n_best;

  size_t m = orc_select_within(cloud, n, best_coeff, threshold, prm->dot_order, inliers);
  tr.n_inliers_raw = (int)m;
  memcpy(coeff, best_coeff, sizeof(best_coeff));
  if (prm->optimize_coefficients) {
    float refined[4];
    if (prm->refit_mode == ORC_REFIT_FIXED) {
      int s = scale_exp_or_min == INT_MIN ? orc_fixed_scale_exp(cloud, n) : scale_exp_or_min;
      tr.scale_exp = s;
      const float pivot[3] = {cloud[best_sample[0]].x, cloud[best_sample[0]].y, cloud[best_sample[0]].z};
      orc_refit_fixed(cloud, inliers, m, pivot, s, best_coeff, refined, NULL);
    } else {
      orc_refit_pcl_float(cloud, inliers, m, best_coeff, refined);
    }
    memcpy(coeff, refined, sizeof(refined));
    /* "Refine inliers": selectWithinDistance(coeff_refined, threshold_, inliers.indices) */
    m = orc_select_within(cloud, n, refined, threshold, prm->dot_order, inliers);
  }
  *n_inliers = m;
  tr.n_inliers = (int)m;
  if (trace) *trace = tr;
  return 1;
}

/* =========================================================================================
 * segment + ExtractIndices peel (filters/impl/extract_indices.hpp): positive = points at the inlier
 * indices, negative = the others in original order.  Loop:
 *   while (planes < max_planes) { segment; if (|inliers| < min_plane_size || |inliers| == 0) break;
 *                                 record; cloud = negative; }
 * ========================================================================================= */
int orc_extract_planes(const orc_point* cloud, size_t n, const orc_params* prm, float* coeffs,
                       int32_t* inlier_cur, int32_t* inlier_orig, size_t idx_cap, size_t* plane_offsets,
                       int* n_planes, orc_point* remaining, size_t* n_remaining, orc_trace* traces) {
  orc_point* cur = (orc_point*)malloc((n ? n : 1) * sizeof(orc_point));
  int32_t* orig = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
  int32_t* inl = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
  if (!cur || !orig || !inl) { free(cur); free(orig); free(inl); return -1; }
  memcpy(cur, cloud, n * sizeof(orc_point));
  for (size_t i = 0; i < n; ++i) orig[i] = (int32_t)i;
  size_t n_cur = n;
  const int scale_exp = orc_fixed_scale_exp(cloud, n); /* extent of the cloud as staged, all rounds */
  int planes = 0;
  int rc = 0;
  plane_offsets[0] = 0;
  while (planes < prm->max_planes) {
    float c[4];
    size_t m = 0;
    orc_trace tr;
    orc_segment(cur, n_cur, prm, scale_exp, c, inl, &m, &tr);
    if (traces) traces[planes] = tr;
    if (m == 0 || m < (size_t)(prm->min_plane_size > 0 ? prm->min_plane_size : 0)) break;
    if (plane_offsets[planes] + m > idx_cap) { rc = -2; break; }
    memcpy(coeffs + 4 * planes, c, sizeof(c));
    size_t off = plane_offsets[planes];
    for (size_t k = 0; k < m; ++k) {
      if (inlier_cur) inlier_cur[off + k] = inl[k];
      if (inlier_orig) inlier_orig[off + k] = orig[inl[k]];
    }
    plane_offsets[planes + 1] = off + m;
    ++planes;
    /* negative extraction, order preserved */
    size_t w = 0, j = 0;
    for (size_t i = 0; i < n_cur; ++i) {
      if (j < m && (size_t)inl[j] == i) { ++j; continue; }
      cur[w] = cur[i];
      orig[w] = orig[i];
      ++w;
    }
    n_cur = w;
  }
  *n_planes = planes;
  if (remaining) memcpy(remaining, cur, n_cur * sizeof(orc_point));
  if (n_remaining) *n_remaining = n_cur;
  free(cur); free(orig); free(inl);
  return rc;
}

/* =========================================================================================
 * Normal estimation: pcl::NormalEstimationOMP<PointXYZ, Normal> with setRadiusSearch(r), the stage in front of
 * the reference's detector (Dialog/PlaneDetect.h:515-545, r_for_estimate_normal = 0.5 in Dialog/config.txt:4).
 * PCL 1.8 features/impl/normal_3d_omp.hpp + normal_3d.h + kdtree/impl/kdtree_flann.hpp (PARITY UNPINNED, see header):
 *   neighbours of p_i : the finite points p_j with L2_Simple distance ((dx*dx + dy*dy) + dz*dz, FP32, dx = p_i.x - p_j.x)
 *                       strictly below (float)(r * r); p_i itself is one of them
 *   fewer than 3 neighbours or a non-finite p_i -> NaN normal and curvature
 *   computePointNormal: centroid + covariance of the neighbours -> eigen33 -> normal = eigenvector of the smallest
 *                       eigenvalue, curvature = |lambda_0 / trace|
 *   flipNormalTowardsViewpoint(p_i, vp): negate when (vp - p_i) . n < 0   (FP32, left to right)
 * ORC_NORMALS_PCL_FLOAT accumulates like PCL (nine float sums over the neighbours in ascending distance order, ties by
 * index — FLANN's order among equal distances is not defined) and runs eigen33 in float.
 * ORC_NORMALS_FIXED is the order-independent form a parallel device reproduces: neighbour coordinates relative to
 * p_i quantised to a 2^-s grid (r * 2^s < 2^18), exact integer moments, 128-bit covariance numerators, eigen33 in
 * double; same neighbour sets, same NaN pattern.
 * ========================================================================================= */
typedef struct { float d; int32_t j; } orc_nb;

static int nb_cmp(const void* a, const void* b) {
  const orc_nb* x = (const orc_nb*)a;
  const orc_nb* y = (const orc_nb*)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return (x->j > y->j) - (x->j < y->j);
}

static inline int finite_pt(const orc_point* p) { return isfinite(p->x) && isfinite(p->y) && isfinite(p->z); }

int orc_normals_scale_exp(double radius) {
  int e;
  (void)frexp(radius, &e); /* radius < 2^e */
  return 18 - e;
}

int orc_estimate_normals(const orc_point* cloud, size_t n, double radius, const float vp[3], int mode, float* out,
                         int32_t* n_neighbors) {
  if (!(radius > 0.0) || !isfinite(radius)) return -1;
  const float r2 = (float)(radius * radius);
  const int s = orc_normals_scale_exp(radius);
  const double scale = ldexp(1.0, s);
  int fail = 0;
#pragma omp parallel
  {
    orc_nb* nb = (orc_nb*)malloc((n ? n : 1) * sizeof(orc_nb));
    if (!nb) {
#pragma omp atomic write
      fail = 1;
    }
#pragma omp for schedule(dynamic, 64)
    for (long long ii = 0; ii < (long long)n; ++ii) {
      const size_t i = (size_t)ii;
      float* o = out + 4 * i;
      o[0] = o[1] = o[2] = o[3] = NAN;
      if (n_neighbors) n_neighbors[i] = 0;
      if (!nb) continue;
      const orc_point* p = &cloud[i];
      if (!finite_pt(p)) continue;
      size_t m = 0;
      for (size_t j = 0; j < n; ++j) {
        const orc_point* q = &cloud[j];
        if (!finite_pt(q)) continue;
        const float dx = p->x - q->x, dy = p->y - q->y, dz = p->z - q->z;
        const float d = (dx * dx + dy * dy) + dz * dz;
        if (d < r2) { nb[m].d = d; nb[m].j = (int32_t)j; ++m; }
      }
      if (n_neighbors) n_neighbors[i] = (int32_t)m;
      if (m < 3) continue;
      float nx, ny, nz, curv;
      if (mode == ORC_NORMALS_PCL_FLOAT) {
        qsort(nb, m, sizeof(orc_nb), nb_cmp);
        float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (size_t k = 0; k < m; ++k) {
          const orc_point* q = &cloud[nb[k].j];
          accu[0] += q->x * q->x; accu[1] += q->x * q->y; accu[2] += q->x * q->z;
          accu[3] += q->y * q->y; accu[4] += q->y * q->z; accu[5] += q->z * q->z;
          accu[6] += q->x; accu[7] += q->y; accu[8] += q->z;
        }
        const float cnt = (float)m;
        for (int k = 0; k < 9; ++k) accu[k] = accu[k] / cnt;
        float cov[9];
        cov[0] = accu[0] - accu[6] * accu[6]; cov[1] = accu[1] - accu[6] * accu[7]; cov[2] = accu[2] - accu[6] * accu[8];
        cov[4] = accu[3] - accu[7] * accu[7]; cov[5] = accu[4] - accu[7] * accu[8]; cov[8] = accu[5] - accu[8] * accu[8];
        cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
        float ev, v[3];
        eigen33_f(cov, &ev, v);
        const float sum = cov[0] + cov[4] + cov[8];
        curv = sum != 0.0f ? fabsf(ev / sum) : 0.0f;
        nx = v[0]; ny = v[1]; nz = v[2];
      } else {
        int64_t S[3] = {0, 0, 0};
        typedef __int128 i128n;
        int64_t Q[6] = {0, 0, 0, 0, 0, 0};
        for (size_t k = 0; k < m; ++k) {
          const orc_point* q = &cloud[nb[k].j];
          const int64_t a = llrint(((double)q->x - (double)p->x) * scale);
          const int64_t b = llrint(((double)q->y - (double)p->y) * scale);
          const int64_t c = llrint(((double)q->z - (double)p->z) * scale);
          S[0] += a; S[1] += b; S[2] += c;
          Q[0] += a * a; Q[1] += a * b; Q[2] += a * c; Q[3] += b * b; Q[4] += b * c; Q[5] += c * c;
        }
        static const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {0, 1, 2, 1, 2, 2};
        double C[6];
        for (int k = 0; k < 6; ++k) C[k] = (double)((i128n)(int64_t)m * (i128n)Q[k] - (i128n)S[A[k]] * (i128n)S[B[k]]);
        const double cov[9] = {C[0], C[1], C[2], C[1], C[3], C[4], C[2], C[4], C[5]};
        double ev, v[3];
        eigen33_d(cov, &ev, v);
        const double sum = C[0] + C[3] + C[5];
        curv = sum != 0.0 ? (float)fabs(ev / sum) : 0.0f;
        nx = (float)v[0]; ny = (float)v[1]; nz = (float)v[2];
      }
      /* flipNormalTowardsViewpoint (normal_3d.h, the float& overload) */
      const float vx = vp[0] - p->x, vy = vp[1] - p->y, vz = vp[2] - p->z;
      const float cos_theta = (vx * nx + vy * ny + vz * nz);
      if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
      o[0] = nx; o[1] = ny; o[2] = nz; o[3] = curv;
    }
    free(nb);
  }
  return fail ? -1 : 0;
}

/* =========================================================================================
 * clusterFilt (Dialog/PlaneDetect.h:1582-1656; reference code): clusters are grown by BFS over kd-tree radius searches
 * (radius_local) and every cluster with at most T_cluster_num points is dropped from source_cloud ("indices.size() <=
 * T_cluster_num").  The clusters are the connected components of the radius graph, whatever the seed order; an edge is
 * FLANN's predicate (FP32 L2_Simple distance strictly below (float)(r * r)).  keep[i] = 0 for dropped points.
 * Deviation noted in include/plane_ransac.h: the reference skips the first search hit as "the query itself", which can
 * split exact duplicates off; here they are connected.  Non-finite points are left alone (keep = 1).
 * ========================================================================================= */
static int32_t uf_root(int32_t* parent, int32_t i) {
  while (parent[i] != i) {
    parent[i] = parent[parent[i]];
    i = parent[i];
  }
  return i;
}

int orc_cluster_filter(const orc_point* cloud, size_t n, double radius, int max_small_cluster, uint8_t* keep) {
  if (!(radius > 0.0) || !isfinite(radius)) return -1;
  const float r2 = (float)(radius * radius);
  int32_t* parent = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
  int32_t* size = (int32_t*)calloc(n ? n : 1, sizeof(int32_t));
  if (!parent || !size) { free(parent); free(size); return -1; }
  for (size_t i = 0; i < n; ++i) parent[i] = (int32_t)i;
  for (size_t i = 0; i < n; ++i) {
    const orc_point* p = &cloud[i];
    if (!finite_pt(p)) continue;
    for (size_t j = 0; j < i; ++j) {
      const orc_point* q = &cloud[j];
      if (!finite_pt(q)) continue;
      const float dx = p->x - q->x, dy = p->y - q->y, dz = p->z - q->z;
      const float d = (dx * dx + dy * dy) + dz * dz;
      if (d < r2) {
        const int32_t a = uf_root(parent, (int32_t)i), b = uf_root(parent, (int32_t)j);
        if (a != b) parent[a > b ? a : b] = a > b ? b : a;
      }
    }
  }
  for (size_t i = 0; i < n; ++i) size[uf_root(parent, (int32_t)i)]++;
  for (size_t i = 0; i < n; ++i)
    keep[i] = (finite_pt(&cloud[i]) && size[uf_root(parent, (int32_t)i)] <= max_small_cluster) ? 0 : 1;
  free(parent);
  free(size);
  return 0;
}

/* =========================================================================================
 * SACMODEL_LINE as the reference calls it (Dialog/SimplifyVerticesSize.cpp:64-67,87,122,146):
 * pcl::SACSegmentation<PointXYZ>, SACMODEL_LINE, SAC_RANSAC, setDistanceThreshold(FLT_MAX), everything else at its
 * default (max_iterations 50, probability 0.99, optimize_coefficients true), on a few contour vertices.  Restated from
 * PCL 1.8 sac_model_line.hpp; it runs through orc_ransac_compute_model and orc_sampler above — a second consumer of
 * the sampler and the loop, with the reference's own call pattern: every point is an inlier of the first good sample,
 * so w = 1, k collapses and the loop ends after one iteration.
 *   isSampleGood            all three coordinates of the two points differ (PCL 1.8 joins the tests with &&)
 *   computeModelCoefficients  (p0, normalised p1 - p0)
 *   countWithinDistance     |(p0 - p) x dir|^2 < threshold^2, the square taken in double
 *   optimizeModelCoefficients  <= 2 inliers: unchanged; else point = centroid, direction = principal axis of the
 *                           covariance.  CHOICE: PCL gets the axis from eigen33(cov, evals) +
 *                           computeCorrespondingEigenVector in FP32; here a cyclic Jacobi in double — the test asserts
 *                           agreement with an independent float64 PCA to 1e-5, not bits.
 * ========================================================================================= */
static int line_good_cb(const orc_point* cloud, const int32_t* idx) {
  const orc_point *a = &cloud[idx[0]], *b = &cloud[idx[1]];
  return (a->x != b->x) && (a->y != b->y) && (a->z != b->z);
}
static int line_model_cb(const orc_point* cloud, const int32_t* idx, float* c) {
  const orc_point *a = &cloud[idx[0]], *b = &cloud[idx[1]];
  c[0] = a->x; c[1] = a->y; c[2] = a->z;
  float dx = b->x - a->x, dy = b->y - a->y, dz = b->z - a->z;
  float nrm = sqrtf(dx * dx + (dy * dy + dz * dz));
  c[3] = dx / nrm; c[4] = dy / nrm; c[5] = dz / nrm;
  return 1;
}
static int64_t line_count_cb(const orc_point* cloud, size_t n, const float* c, double t, int dot_order) {
  (void)dot_order;
  const double sqr_t = t * t;
  int64_t cnt = 0;
  for (size_t i = 0; i < n; ++i) {
    const float vx = c[0] - cloud[i].x, vy = c[1] - cloud[i].y, vz = c[2] - cloud[i].z;
    const float cx = vy * c[5] - vz * c[4], cy = vz * c[3] - vx * c[5], cz = vx * c[4] - vy * c[3];
    const float d2 = cx * cx + (cy * cy + cz * cz);
    if ((double)d2 < sqr_t) ++cnt;
  }
  return cnt;
}

int orc_segment_line(const orc_point* cloud, size_t n, const orc_params* prm, float coeff[6], int32_t* inliers, size_t* n_inliers,
                     orc_trace* trace) {
  orc_trace tr;
  memset(&tr, 0, sizeof(tr));
  *n_inliers = 0;
  for (int i = 0; i < 6; ++i) coeff[i] = 0.0f;
  const orc_sac_model line = {2, line_good_cb, line_model_cb, line_count_cb, 6};
  orc_sac_result res;
  if (orc_ransac_compute_model(cloud, n, prm, &line, &res)) return -1;
  tr.iterations = res.iterations;
  tr.draws = res.draws;
  tr.skipped = res.skipped;
  tr.ok = res.have_model;
  if (!res.have_model) { if (trace) *trace = tr; return 0; }
  tr.best_sample[0] = res.best_sample[0];
  tr.best_sample[1] = res.best_sample[1];
  tr.best_count = res.n_best;
  memcpy(tr.raw_coeff, res.best_coeff, 4 * sizeof(float));
  /* selectWithinDistance */
  const double sqr_t = prm->distance_threshold * prm->distance_threshold;
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    const float* c = res.best_coeff;
    const float vx = c[0] - cloud[i].x, vy = c[1] - cloud[i].y, vz = c[2] - cloud[i].z;
    const float cx = vy * c[5] - vz * c[4], cy = vz * c[3] - vx * c[5], cz = vx * c[4] - vy * c[3];
    if ((double)(cx * cx + (cy * cy + cz * cz)) < sqr_t) inliers[m++] = (int32_t)i;
  }
  memcpy(coeff, res.best_coeff, 6 * sizeof(float));
  if (prm->optimize_coefficients && m > 2) {
    double mean[3] = {0, 0, 0};
    for (size_t k = 0; k < m; ++k) { mean[0] += cloud[inliers[k]].x; mean[1] += cloud[inliers[k]].y; mean[2] += cloud[inliers[k]].z; }
    for (int a = 0; a < 3; ++a) mean[a] /= (double)m;
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (size_t k = 0; k < m; ++k) {
      const double d[3] = {cloud[inliers[k]].x - mean[0], cloud[inliers[k]].y - mean[1], cloud[inliers[k]].z - mean[2]};
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) A[a][b] += d[a] * d[b];
    }
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 60; ++sweep) {
      for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double th = 0.5 * atan2(2.0 * A[p][q], A[q][q] - A[p][p]);
        const double cs = cos(th), sn = sin(th);
        for (int r = 0; r < 3; ++r) { const double ap = A[r][p], aq = A[r][q]; A[r][p] = cs * ap - sn * aq; A[r][q] = sn * ap + cs * aq; }
        for (int r = 0; r < 3; ++r) { const double ap = A[p][r], aq = A[q][r]; A[p][r] = cs * ap - sn * aq; A[q][r] = sn * ap + cs * aq; }
        for (int r = 0; r < 3; ++r) { const double vp = V[r][p], vq = V[r][q]; V[r][p] = cs * vp - sn * vq; V[r][q] = sn * vp + cs * vq; }
      }
    }
    int big = 0;
    if (A[1][1] > A[big][big]) big = 1;
    if (A[2][2] > A[big][big]) big = 2;
    coeff[0] = (float)mean[0]; coeff[1] = (float)mean[1]; coeff[2] = (float)mean[2];
    coeff[3] = (float)V[0][big]; coeff[4] = (float)V[1][big]; coeff[5] = (float)V[2][big];
    /* "Refine inliers": selectWithinDistance with the optimised line */
    m = 0;
    for (size_t i = 0; i < n; ++i) {
      const float vx = coeff[0] - cloud[i].x, vy = coeff[1] - cloud[i].y, vz = coeff[2] - cloud[i].z;
      const float cx = vy * coeff[5] - vz * coeff[4], cy = vz * coeff[3] - vx * coeff[5], cz = vx * coeff[4] - vy * coeff[3];
      if ((double)(cx * cx + (cy * cy + cz * cz)) < sqr_t) inliers[m++] = (int32_t)i;
    }
  }
  *n_inliers = m;
  tr.n_inliers = (int)m;
  if (trace) *trace = tr;
  return 1;
}

/* =========================================================================================
 * clusterFilt as the reference writes it (Dialog/PlaneDetect.h:1598-1634), literally: seeds in index order, BFS over
 * radius searches whose hits are sorted by distance (pcl::KdTreeFLANN::radiusSearch, sorted results), the FIRST hit of
 * every search skipped as "the query itself" (for (i = 1; ...), :1623).  FLANN does not define the order of hits at
 * equal distance; ties_by_index = 1 sorts them by descending index (an exact duplicate with a larger index then comes
 * before the query and is the one skipped), 0 puts the query first (the skip then always hits the query).  Exists to ENUMERATE
 * where that skip makes the reference differ from the connected-components definition orc_cluster_filter and the
 * device use (tests/test_normals.py): only isolated groups of exact duplicates.
 * ========================================================================================= */
typedef struct { float d; int32_t i; int self; } orc_hit;
static int hit_cmp_self_first(const void* a, const void* b) {
  const orc_hit *x = (const orc_hit*)a, *y = (const orc_hit*)b;
  if (x->d != y->d) return x->d < y->d ? -1 : 1;
  if (x->self != y->self) return x->self ? -1 : 1;
  return (x->i > y->i) - (x->i < y->i);
}
static int hit_cmp_by_index(const void* a, const void* b) {
  const orc_hit *x = (const orc_hit*)a, *y = (const orc_hit*)b;
  if (x->d != y->d) return x->d < y->d ? -1 : 1;
  return (y->i > x->i) - (y->i < x->i); /* descending: a copy with a larger index precedes the query */
}

int orc_cluster_filter_reference_bfs(const orc_point* cloud, size_t n, double radius, int max_small_cluster, int ties_by_index,
                                     uint8_t* keep) {
  if (!(radius > 0.0) || !isfinite(radius)) return -1;
  const float r2 = (float)(radius * radius);
  uint8_t* processed = (uint8_t*)calloc(n ? n : 1, 1);
  int32_t* queue = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
  int32_t* members = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
  orc_hit* hits = (orc_hit*)malloc((n ? n : 1) * sizeof(orc_hit));
  if (!processed || !queue || !members || !hits) { free(processed); free(queue); free(members); free(hits); return -1; }
  for (size_t i = 0; i < n; ++i) keep[i] = 1;
  for (size_t seed = 0; seed < n; ++seed) {
    if (processed[seed] || !finite_pt(&cloud[seed])) continue;
    size_t qh = 0, qt = 0, nm = 0;
    queue[qt++] = (int32_t)seed;
    members[nm++] = (int32_t)seed;
    while (qh < qt) {
      const int32_t cur = queue[qh++];
      processed[cur] = 1;
      size_t nh = 0;
      for (size_t j = 0; j < n; ++j) {
        if (!finite_pt(&cloud[j])) continue;
        const float dx = cloud[cur].x - cloud[j].x, dy = cloud[cur].y - cloud[j].y, dz = cloud[cur].z - cloud[j].z;
        const float d = (dx * dx + dy * dy) + dz * dz;
        if (d < r2) { hits[nh].d = d; hits[nh].i = (int32_t)j; hits[nh].self = (int32_t)j == cur; ++nh; }
      }
      qsort(hits, nh, sizeof(orc_hit), ties_by_index ? hit_cmp_by_index : hit_cmp_self_first);
      for (size_t h = 1; h < nh; ++h) {
        const int32_t j = hits[h].i;
        if (processed[j]) continue;
        queue[qt++] = j;
        members[nm++] = j;
        processed[j] = 1;
      }
    }
    if ((long long)nm <= (long long)max_small_cluster)
      for (size_t k = 0; k < nm; ++k) keep[members[k]] = 0;
  }
  free(processed); free(queue); free(members); free(hits);
  return 0;
}
