// pr_math.h — the closed-form least-squares plane (pcl::computeMeanAndCovarianceMatrix + pcl::eigen33 behind
// pcl::computePointNormal, Dialog/PlaneDetect.h:1084,1130,1386,1485) from exact integer moments, written so that the
// SAME source runs on the host (g++, -ffp-contract=off) and inside a kernel (nvcc, -fmad=false) and gives the same bits:
// only IEEE-754 double +, -, *, /, sqrt and integer operations are used.  PCL's computeRoots calls atan2 / cos / sin;
// those come from the portable implementations below (fdlibm's algorithms: argument reduction by table for atan,
// minimax polynomials on [-pi/4, pi/4] for sin and cos), accurate to < 1 ulp, instead of a platform libm, because
// glibc's and CUDA's functions differ in the last bit and the peel loop runs without the host in it.
// The CPU checker under tests/ keeps libm; tests/test_host_logic.py checks that both give the same FP32 coefficients.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define PR_HD __host__ __device__ __forceinline__
#else
#define PR_HD inline
#endif

namespace pr {

PR_HD double pm_from_bits(uint64_t b) {
  union { uint64_t u; double d; } v;
  v.u = b;
  return v.d;
}
PR_HD uint64_t pm_bits(double d) {
  union { uint64_t u; double d; } v;
  v.d = d;
  return v.u;
}
PR_HD double pm_fabs(double x) { return pm_from_bits(pm_bits(x) & 0x7FFFFFFFFFFFFFFFull); }
// 2^e for -1022 <= e <= 1023 (exact)
PR_HD double pm_pow2(int e) { return pm_from_bits((uint64_t)(e + 1023) << 52); }
PR_HD double pm_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(x);
#else
  return __builtin_sqrt(x);
#endif
}
PR_HD int pm_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}

// (double)v, correctly rounded (round to nearest even), for any 128-bit integer: the top 64 bits with a sticky bit
// are converted by the hardware's exact u64 -> f64 rounding and scaled by a power of two.
PR_HD double pm_i128_to_double(__int128 v) {
  const bool neg = v < 0;
  const unsigned __int128 u = neg ? (unsigned __int128)0 - (unsigned __int128)v : (unsigned __int128)v;
  const uint64_t hi = (uint64_t)(u >> 64), lo = (uint64_t)u;
  double d;
  if (hi == 0) {
    d = (double)lo;
  } else {
    const int shift = 64 - pm_clz64(hi);  // 1..64 bits above the low word
    uint64_t top = (uint64_t)(u >> shift);
    const unsigned __int128 below = u & ((((unsigned __int128)1) << shift) - 1);
    if (below != 0) top |= 1ull;  // sticky: far below the rounding position of a 64 -> 53 bit conversion
    d = (double)top * pm_pow2(shift);
  }
  return neg ? -d : d;
}

// ---- sin / cos on [-pi/4, pi/4] (fdlibm __kernel_sin / __kernel_cos; y is the tail of the argument) -----------------
PR_HD double pm_ksin(double x, double y, bool have_tail) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = x * x;
  const double v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  if (!have_tail) return x + v * (S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

PR_HD double pm_kcos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = x * x;
  const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  const double ax = pm_fabs(x);
  if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
  double qx;
  if (ax > 0.78125) {
    qx = 0.28125;
  } else {
    // x / 4 with the low word cleared
    qx = pm_from_bits(((pm_bits(ax) >> 32) - 0x00200000ull) << 32);
  }
  const double hz = 0.5 * z - qx;
  const double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}

// sin and cos of theta in [0, pi/2]
PR_HD void pm_sincos(double theta, double* s, double* c) {
  const double pio4 = 7.85398163397448278999e-01;
  const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
  if (!(theta > pio4)) {  // also NaN: propagates through the polynomials
    *s = pm_ksin(theta, 0.0, false);
    *c = pm_kcos(theta, 0.0);
    return;
  }
  const double r = pio2_hi - theta;  // exact (Sterbenz) for theta in [pi/4, pi/2]
  *s = pm_kcos(r, pio2_lo);
  *c = pm_ksin(r, pio2_lo, true);
}

// atan(x) for x >= 0 (fdlibm s_atan.c)
PR_HD double pm_atan_pos(double x) {
  const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                            1.57079632679489655800e+00};
  const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
                            6.12323399573676603587e-17};
  const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                         -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                         6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                         -3.65315727442169155270e-02, 1.62858201153657823623e-02};
  if (x != x) return x;
  if (x >= 7.378697629483821e19) return atanhi[3] + atanlo[3];  // 2^66
  int id;
  if (x < 0.4375) {
    if (x < 1.862645149230957e-09) return x;  // 2^-29
    id = -1;
  } else if (x < 1.1875) {
    if (x < 0.6875) {
      id = 0;
      x = (2.0 * x - 1.0) / (2.0 + x);
    } else {
      id = 1;
      x = (x - 1.0) / (x + 1.0);
    }
  } else if (x < 2.4375) {
    id = 2;
    x = (x - 1.5) / (1.0 + 1.5 * x);
  } else {
    id = 3;
    x = -1.0 / x;
  }
  const double z = x * x;
  const double w = z * z;
  const double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  const double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  return atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
}

// atan2(y, x) for y >= 0 (computeRoots calls it with y = sqrt(-q)); result in [0, pi]
PR_HD double pm_atan2_ypos(double y, double x) {
  const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16;
  const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
  if (x != x || y != y) return x + y;
  const bool xneg = (pm_bits(x) >> 63) != 0;
  if (y == 0.0) return xneg ? pi : 0.0;
  if (x == 0.0) return pio2_hi + 0.5 * pio2_lo;
  const double ax = pm_fabs(x);
  const int ey = (int)((pm_bits(y) >> 52) & 0x7FF), ex = (int)((pm_bits(ax) >> 52) & 0x7FF);
  const int k = ey - ex;
  double z;
  if (k > 60) z = pio2_hi + 0.5 * pio2_lo;
  else if (xneg && k < -60) z = 0.0;
  else z = pm_atan_pos(y / ax);
  if (!xneg) return z;
  return pi - (z - pi_lo);
}

// ---- pcl::eigen33 / computeRoots / computeRoots2 (common/impl/eigen.hpp) in double ---------------------------------
PR_HD void pm_roots2(double b, double c, double roots[3]) {
  roots[0] = 0.0;
  double d = b * b - 4.0 * c;
  if (d < 0.0) d = 0.0;
  const double sd = pm_sqrt(d);
  roots[2] = 0.5 * (b + sd);
  roots[1] = 0.5 * (b - sd);
}

PR_HD void pm_roots3(const double m[9], double roots[3]) {
  const double kEps = 2.2204460492503131e-16;  // DBL_EPSILON
  const double c0 = m[0] * m[4] * m[8] + 2.0 * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] -
                    m[8] * m[1] * m[1];
  const double c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  const double c2 = m[0] + m[4] + m[8];
  if (pm_fabs(c0) < kEps) {
    pm_roots2(c2, c1, roots);
    return;
  }
  const double s_inv3 = 1.0 / 3.0;
  const double s_sqrt3 = pm_sqrt(3.0);
  const double c2_over_3 = c2 * s_inv3;
  double a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0) a_over_3 = 0.0;
  const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
  double q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0) q = 0.0;
  const double rho = pm_sqrt(-a_over_3);
  const double theta = pm_atan2_ypos(pm_sqrt(-q), half_b) * s_inv3;  // in [0, pi/3]
  double cos_theta, sin_theta;
  pm_sincos(theta, &sin_theta, &cos_theta);
  roots[0] = c2_over_3 + 2.0 * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  double tmp;
  if (roots[0] >= roots[1]) { tmp = roots[0]; roots[0] = roots[1]; roots[1] = tmp; }
  if (roots[1] >= roots[2]) {
    tmp = roots[1]; roots[1] = roots[2]; roots[2] = tmp;
    if (roots[0] >= roots[1]) { tmp = roots[0]; roots[0] = roots[1]; roots[1] = tmp; }
  }
  if (roots[0] <= 0) pm_roots2(c2, c1, roots);
}

PR_HD void pm_smallest_eigenvector(const double mat[9], double vec[3]) {
  const double kMin = 2.2250738585072014e-308;  // DBL_MIN
  double scale = 0.0;
  for (int i = 0; i < 9; ++i) {
    const double a = pm_fabs(mat[i]);
    if (a > scale) scale = a;
  }
  if (scale <= kMin) scale = 1.0;
  double s[9];
  for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;
  double ev[3];
  pm_roots3(s, ev);
  s[0] -= ev[0];
  s[4] -= ev[0];
  s[8] -= ev[0];
  const double v1[3] = {s[1] * s[5] - s[2] * s[4], s[2] * s[3] - s[0] * s[5], s[0] * s[4] - s[1] * s[3]};
  const double v2[3] = {s[1] * s[8] - s[2] * s[7], s[2] * s[6] - s[0] * s[8], s[0] * s[7] - s[1] * s[6]};
  const double v3[3] = {s[4] * s[8] - s[5] * s[7], s[5] * s[6] - s[3] * s[8], s[3] * s[7] - s[4] * s[6]};
  const double len1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  const double len2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  const double len3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  double bx, by, bz, len;
  if (len1 >= len2 && len1 >= len3) { bx = v1[0]; by = v1[1]; bz = v1[2]; len = len1; }
  else if (len2 >= len1 && len2 >= len3) { bx = v2[0]; by = v2[1]; bz = v2[2]; len = len2; }
  else { bx = v3[0]; by = v3[1]; bz = v3[2]; len = len3; }
  const double nrm = pm_sqrt(len);
  vec[0] = bx / nrm;
  vec[1] = by / nrm;
  vec[2] = bz / nrm;
}

// Least-squares plane from the 16 integer moments of the refit pass (DESIGN.md §3 "Refit"): m = {n, Sx, Sy, Sz,
// then (hi, lo) of Sxx, Sxy, Sxz, Syy, Syz, Szz with S_ab = hi * 2^32 + lo} on the 2^-scale_exp grid about `pivot`.
// Returns false (coeff untouched) when fewer than 4 points contributed (PCL keeps the sample's model then).
PR_HD bool pm_plane_from_moments(const long long m[16], const float pivot[3], int scale_exp, float coeff[4]) {
  typedef __int128 i128;
  const long long n = m[0];
  if (n < 4) return false;
  const i128 S[3] = {(i128)m[1], (i128)m[2], (i128)m[3]};
  i128 Sab[6];
  for (int k = 0; k < 6; ++k) Sab[k] = (i128)m[4 + 2 * k] * ((i128)1 << 32) + (i128)m[5 + 2 * k];
  const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {0, 1, 2, 1, 2, 2};
  double C[6];
  for (int k = 0; k < 6; ++k) C[k] = pm_i128_to_double((i128)n * Sab[k] - S[A[k]] * S[B[k]]);
  const double cov[9] = {C[0], C[1], C[2], C[1], C[3], C[4], C[2], C[4], C[5]};
  double v[3];
  pm_smallest_eigenvector(cov, v);
  const double inv = pm_pow2(-scale_exp);
  const double cx = (double)pivot[0] + ((double)m[1] / (double)n) * inv;
  const double cy = (double)pivot[1] + ((double)m[2] / (double)n) * inv;
  const double cz = (double)pivot[2] + ((double)m[3] / (double)n) * inv;
  const double d = -((v[0] * cx + v[1] * cy) + v[2] * cz);
  coeff[0] = (float)v[0];
  coeff[1] = (float)v[1];
  coeff[2] = (float)v[2];
  coeff[3] = (float)d;
  return true;
}

}  // namespace pr
