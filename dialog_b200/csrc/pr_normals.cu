// pr_normals.cu — point normals by radius PCA: pcl::NormalEstimationOMP<PointXYZ, Normal> with setRadiusSearch(r),
// the stage in front of the reference's detector (Dialog/PlaneDetect.h:515-545), on sm_100a.
//
//   N1  normals_cell_keys_kernel   uniform grid of cell size r (1 + 1e-6): 64-bit linear cell key per finite point
//   --  cub::DeviceRadixSort       (key, index) pairs -> points in cell order; x-adjacent cells are adjacent keys
//   N2  normals_gather_kernel      coordinates in sorted order
//   N3  normals_kernel             thread per point: 9 rows of 3 x-adjacent cells = 9 contiguous ranges found by binary
//                                  search; FLANN's L2_Simple test d2 < (float)(r*r) in FP32 on every candidate; exact
//                                  integer moments of the neighbours relative to the point on a 2^-s grid; 128-bit
//                                  covariance numerators; PCL's eigen33 closed form in double; curvature; flip to the
//                                  viewpoint.  Neighbour sets, counts and the NaN pattern equal PCL's exactly; the
//                                  moments are order independent (integer sums), so the result does not depend on the
//                                  traversal.
#include "pr_kernels.h"

#include <cfloat>
#include <math_constants.h>

#include <cub/device/device_radix_sort.cuh>

namespace pr {

namespace {

__global__ void __launch_bounds__(256) normals_cell_keys_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                const float* __restrict__ z, size_t n, NormalsGrid g,
                                                                unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float px = x[i], py = y[i], pz = z[i];
    unsigned long long k = g.no_cell;  // non-finite points: one past the last cell, they sort to the end
    if (isfinite(px) && isfinite(py) && isfinite(pz)) {
      long long cx = (long long)floor(((double)px - g.lo[0]) * g.inv_h);
      long long cy = (long long)floor(((double)py - g.lo[1]) * g.inv_h);
      long long cz = (long long)floor(((double)pz - g.lo[2]) * g.inv_h);
      cx = min(max(cx, 0ll), g.dim[0] - 1);
      cy = min(max(cy, 0ll), g.dim[1] - 1);
      cz = min(max(cz, 0ll), g.dim[2] - 1);
      k = (unsigned long long)((cz * g.dim[1] + cy) * g.dim[0] + cx);
    }
    keys[i] = k;
    idx[i] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(256) normals_gather_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ z, const uint32_t* __restrict__ idx, size_t n,
                                                             float* __restrict__ sx, float* __restrict__ sy, float* __restrict__ sz) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t j = idx[i];
    sx[i] = x[j];
    sy[i] = y[j];
    sz[i] = z[j];
  }
}

// first position in keys[0, n) whose key is >= k
__device__ __forceinline__ size_t lower_bound_key(const unsigned long long* __restrict__ keys, size_t n, unsigned long long k) {
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- pcl::eigen33 / computeRoots / computeRoots2 (common/impl/eigen.hpp) in double ---------------------------------
__device__ void roots2(double b, double c, double roots[3]) {
  roots[0] = 0.0;
  double d = b * b - 4.0 * c;
  if (d < 0.0) d = 0.0;
  const double sd = sqrt(d);
  roots[2] = 0.5 * (b + sd);
  roots[1] = 0.5 * (b - sd);
}

__device__ void roots3(const double m[9], double roots[3]) {
  const double c0 = m[0] * m[4] * m[8] + 2.0 * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
  const double c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  const double c2 = m[0] + m[4] + m[8];
  if (fabs(c0) < DBL_EPSILON) {
    roots2(c2, c1, roots);
    return;
  }
  const double s_inv3 = 1.0 / 3.0;
  const double s_sqrt3 = sqrt(3.0);
  const double c2_over_3 = c2 * s_inv3;
  double a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0) a_over_3 = 0.0;
  const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
  double q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0) q = 0.0;
  const double rho = sqrt(-a_over_3);
  const double theta = atan2(sqrt(-q), half_b) * s_inv3;
  const double cos_theta = cos(theta);
  const double sin_theta = sin(theta);
  roots[0] = c2_over_3 + 2.0 * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  double tmp;
  if (roots[0] >= roots[1]) { tmp = roots[0]; roots[0] = roots[1]; roots[1] = tmp; }
  if (roots[1] >= roots[2]) {
    tmp = roots[1]; roots[1] = roots[2]; roots[2] = tmp;
    if (roots[0] >= roots[1]) { tmp = roots[0]; roots[0] = roots[1]; roots[1] = tmp; }
  }
  if (roots[0] <= 0) roots2(c2, c1, roots);
}

__device__ void eigen33(const double mat[9], double* eigenvalue, double vec[3]) {
  double scale = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) scale = fmax(scale, fabs(mat[i]));
  if (scale <= DBL_MIN) scale = 1.0;
  double s[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;
  double ev[3];
  roots3(s, ev);
  *eigenvalue = ev[0] * scale;
  s[0] -= ev[0];
  s[4] -= ev[0];
  s[8] -= ev[0];
  const double *r0 = s, *r1 = s + 3, *r2 = s + 6;
  const double v1[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
  const double v2[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
  const double v3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  const double len1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  const double len2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  const double len3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const double* best;
  double len;
  if (len1 >= len2 && len1 >= len3) { best = v1; len = len1; }
  else if (len2 >= len1 && len2 >= len3) { best = v2; len = len2; }
  else { best = v3; len = len3; }
  const double nrm = sqrt(len);
  vec[0] = best[0] / nrm;
  vec[1] = best[1] / nrm;
  vec[2] = best[2] / nrm;
}

__global__ void __launch_bounds__(128) normals_kernel(const float* __restrict__ sx, const float* __restrict__ sy,
                                                      const float* __restrict__ sz, const unsigned long long* __restrict__ keys,
                                                      const uint32_t* __restrict__ idx, size_t n, NormalsGrid g, float r2,
                                                      double scale, float vpx, float vpy, float vpz, float4* __restrict__ out,
                                                      int32_t* __restrict__ n_neighbors) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t dst = idx[i];
  float4 res = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
  int cnt = 0;
  const size_t n_finite = lower_bound_key(keys, n, g.no_cell);
  if (i < n_finite) {  // non-finite points carry the largest key and sort to the end
    const float px = sx[i], py = sy[i], pz = sz[i];
    const unsigned long long k = keys[i];
    const long long cx = (long long)(k % (unsigned long long)g.dim[0]);
    const long long cy = (long long)((k / (unsigned long long)g.dim[0]) % (unsigned long long)g.dim[1]);
    const long long cz = (long long)(k / ((unsigned long long)g.dim[0] * (unsigned long long)g.dim[1]));
    const long long x0 = max(cx - 1, 0ll), x1 = min(cx + 1, g.dim[0] - 1);
    long long S0 = 0, S1 = 0, S2 = 0, Q0 = 0, Q1 = 0, Q2 = 0, Q3 = 0, Q4 = 0, Q5 = 0;
    for (long long zz = max(cz - 1, 0ll); zz <= min(cz + 1, g.dim[2] - 1); ++zz) {
      for (long long yy = max(cy - 1, 0ll); yy <= min(cy + 1, g.dim[1] - 1); ++yy) {
        const unsigned long long row = (unsigned long long)((zz * g.dim[1] + yy) * g.dim[0]);
        const size_t b = lower_bound_key(keys, n_finite, row + (unsigned long long)x0);
        const size_t e = lower_bound_key(keys, n_finite, row + (unsigned long long)x1 + 1ull);
        for (size_t j = b; j < e; ++j) {
          const float qx = __ldg(sx + j), qy = __ldg(sy + j), qz = __ldg(sz + j);
          // flann::L2_Simple: result += diff * diff over x, y, z, FP32, query minus data
          const float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
          const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          if (d2 < r2) {
            const long long a = __double2ll_rn(((double)qx - (double)px) * scale);
            const long long bq = __double2ll_rn(((double)qy - (double)py) * scale);
            const long long c = __double2ll_rn(((double)qz - (double)pz) * scale);
            ++cnt;
            S0 += a; S1 += bq; S2 += c;
            Q0 += a * a; Q1 += a * bq; Q2 += a * c; Q3 += bq * bq; Q4 += bq * c; Q5 += c * c;
          }
        }
      }
    }
    if (cnt >= 3) {
      const __int128 m = cnt;
      const long long S[3] = {S0, S1, S2};
      const long long Q[6] = {Q0, Q1, Q2, Q3, Q4, Q5};
      const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {0, 1, 2, 1, 2, 2};
      double C[6];
#pragma unroll
      for (int t = 0; t < 6; ++t) C[t] = (double)(m * (__int128)Q[t] - (__int128)S[A[t]] * (__int128)S[B[t]]);
      const double cov[9] = {C[0], C[1], C[2], C[1], C[3], C[4], C[2], C[4], C[5]};
      double ev, v[3];
      eigen33(cov, &ev, v);
      const double sum = C[0] + C[3] + C[5];
      const float curv = sum != 0.0 ? (float)fabs(ev / sum) : 0.0f;
      float nx = (float)v[0], ny = (float)v[1], nz = (float)v[2];
      // flipNormalTowardsViewpoint (normal_3d.h)
      const float vx = __fsub_rn(vpx, px), vy = __fsub_rn(vpy, py), vz = __fsub_rn(vpz, pz);
      const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, nx), __fmul_rn(vy, ny)), __fmul_rn(vz, nz));
      if (cos_theta < 0) { nx = -nx; ny = -ny; nz = -nz; }
      res = make_float4(nx, ny, nz, curv);
    }
  }
  out[dst] = res;
  if (n_neighbors) n_neighbors[dst] = cnt;
}

// ---- clusterFilt (Dialog/PlaneDetect.h:1582-1656): connected components of the radius graph ---------------------
// The reference grows clusters by BFS over kd-tree radius searches and drops every cluster with at most T_cluster_num
// points.  Components do not depend on the traversal, so a lock-free union-find over the same cell-sorted points gives
// the same partition: every point hooks to each earlier (sorted order) point within the radius.
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t i) {
  volatile uint32_t* vp = parent;  // other threads hook and compress concurrently: always read memory
  uint32_t p = vp[i];
  while (p != i) {  // path halving (every value ever stored in parent[x] is an ancestor of x)
    const uint32_t gp = vp[p];
    if (gp != p) vp[i] = gp;
    i = p;
    p = gp;
  }
  return i;
}

__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const uint32_t t = a; a = b; b = t; }  // hook the larger root under the smaller
    const uint32_t old = atomicCAS(&parent[a], a, b);
    if (old == a) return;
  }
}

// After the union kernel has finished the forest is final: roots are looked up without writing (a path-compressing
// find running beside another thread's final store could leave a non-root behind).
__device__ __forceinline__ uint32_t uf_root_readonly(const uint32_t* __restrict__ parent, uint32_t i) {
  uint32_t p = parent[i];
  while (p != i) {
    i = p;
    p = parent[i];
  }
  return i;
}

__global__ void __launch_bounds__(256) cluster_init_kernel(uint32_t* __restrict__ parent, uint32_t* __restrict__ sizes, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    parent[i] = (uint32_t)i;
    sizes[i] = 0u;
  }
}

__global__ void __launch_bounds__(128) cluster_union_kernel(const float* __restrict__ sx, const float* __restrict__ sy,
                                                            const float* __restrict__ sz, const unsigned long long* __restrict__ keys,
                                                            size_t n, NormalsGrid g, float r2, uint32_t* __restrict__ parent) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t n_finite = lower_bound_key(keys, n, g.no_cell);
  if (i >= n_finite) return;
  const float px = sx[i], py = sy[i], pz = sz[i];
  const unsigned long long k = keys[i];
  const long long cx = (long long)(k % (unsigned long long)g.dim[0]);
  const long long cy = (long long)((k / (unsigned long long)g.dim[0]) % (unsigned long long)g.dim[1]);
  const long long cz = (long long)(k / ((unsigned long long)g.dim[0] * (unsigned long long)g.dim[1]));
  const long long x0 = max(cx - 1, 0ll), x1 = min(cx + 1, g.dim[0] - 1);
  for (long long zz = max(cz - 1, 0ll); zz <= min(cz + 1, g.dim[2] - 1); ++zz) {
    for (long long yy = max(cy - 1, 0ll); yy <= min(cy + 1, g.dim[1] - 1); ++yy) {
      const unsigned long long row = (unsigned long long)((zz * g.dim[1] + yy) * g.dim[0]);
      const size_t b = lower_bound_key(keys, n_finite, row + (unsigned long long)x0);
      size_t e = lower_bound_key(keys, n_finite, row + (unsigned long long)x1 + 1ull);
      if (e > i) e = i;  // each pair once: only earlier points
      for (size_t j = b; j < e; ++j) {
        const float dx = __fsub_rn(px, __ldg(sx + j)), dy = __fsub_rn(py, __ldg(sy + j)), dz = __fsub_rn(pz, __ldg(sz + j));
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d2 < r2 && uf_find(parent, (uint32_t)i) != uf_find(parent, (uint32_t)j)) uf_union(parent, (uint32_t)i, (uint32_t)j);
      }
    }
  }
}

__global__ void __launch_bounds__(256) cluster_count_kernel(const uint32_t* __restrict__ parent, uint32_t* __restrict__ sizes, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(&sizes[uf_root_readonly(parent, (uint32_t)i)], 1u);
}

// flags[original index] = 1 for every point to drop: clusters with at most max_small points; non-finite points form no
// cluster in the reference's kd-tree and are kept as they are (keep_nonfinite) — they never reach clusterFilt there.
__global__ void __launch_bounds__(256) cluster_flag_kernel(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ sizes,
                                                           const uint32_t* __restrict__ idx, const unsigned long long* __restrict__ keys,
                                                           size_t n, unsigned long long no_cell, uint32_t max_small,
                                                           uint32_t* __restrict__ flags) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool finite = keys[i] != no_cell;
    flags[idx[i]] = (finite && sizes[uf_root_readonly(parent, (uint32_t)i)] <= max_small) ? 1u : 0u;
  }
}

}  // namespace

void launch_cluster_flags(const float* sorted_xyz, const unsigned long long* sorted_keys, const uint32_t* sorted_idx, size_t n,
                          const NormalsGrid& g, float r2, uint32_t max_small, uint32_t* parent, uint32_t* sizes, uint32_t* flags,
                          cudaStream_t s) {
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cluster_init_kernel<<<(unsigned)blocks, 256, 0, s>>>(parent, sizes, n);
  cluster_union_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(sorted_xyz, sorted_xyz + n, sorted_xyz + 2 * n, sorted_keys, n, g, r2, parent);
  cluster_count_kernel<<<(unsigned)blocks, 256, 0, s>>>(parent, sizes, n);
  cluster_flag_kernel<<<(unsigned)blocks, 256, 0, s>>>(parent, sizes, sorted_idx, sorted_keys, n, g.no_cell, max_small, flags);
}

size_t normals_sort_temp_bytes(size_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
  return bytes;
}

void launch_normals_sort(CloudView cloud, size_t n, const NormalsGrid& g, int key_bits, unsigned long long* keys /* 2n */,
                         uint32_t* idx /* 2n */, void* temp, size_t temp_bytes, float* sorted_xyz /* 3n */, cudaStream_t s) {
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  normals_cell_keys_kernel<<<(unsigned)blocks, 256, 0, s>>>(cloud.x, cloud.y, cloud.z, n, g, keys, idx);
  cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys + n, idx, idx + n, (int)n, 0, key_bits, s);
  normals_gather_kernel<<<(unsigned)blocks, 256, 0, s>>>(cloud.x, cloud.y, cloud.z, idx + n, n, sorted_xyz, sorted_xyz + n,
                                                       sorted_xyz + 2 * n);
}

void launch_normals(const float* sorted_xyz, const unsigned long long* sorted_keys, const uint32_t* sorted_idx, size_t n,
                    const NormalsGrid& g, float r2, double scale, const float vp[3], float4* out, int32_t* n_neighbors,
                    cudaStream_t s) {
  if (n == 0) return;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  normals_kernel<<<blocks, 128, 0, s>>>(sorted_xyz, sorted_xyz + n, sorted_xyz + 2 * n, sorted_keys, sorted_idx, n, g, r2,
                                        scale, vp[0], vp[1], vp[2], out, n_neighbors);
}

}  // namespace pr
