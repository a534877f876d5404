// pr_reabsorb.cu — the reference's postProcessPlanes re-absorption pass (Dialog/PlaneDetect.h:1454-1580) on sm_100a.
//
// Every still-unclaimed point is tested against every plane polygon with the reference's isPointInPoly
// (:1891-1955): point-to-plane distance by projPoint2Plane / distP2P (:1442-1448, 203-207, 2019-2023), then ten
// rays in the plane, each perpendicular to a randomly drawn border edge, intersected with every border edge by
// isBothLineSegsIntersect (:1957-2016); the point is inside when at least five rays cross the border an odd
// number of times.  A point is claimed by every plane that contains it (the reference does not break).
//
//   R1  reabsorb_filter_kernel   HBM-bound: 12 B per point, all P planes per load; emits the (point, plane)
//                                 pairs with dist <= T (non-strict, as the reference) — typically a few per cent
//   R2  reabsorb_poly_kernel      FP32-bound: one warp per pair, the 10 x |border| segment tests spread over the
//                                 lanes, ray-crossing parities combined with one warp XOR-reduction
//   R3  sort of the (plane, index) keys -> per-plane ascending index lists; K5 compaction of the unclaimed points
//
// All arithmetic is the reference's, operation for operation, FP32 with separately rounded products and sums
// (this file is compiled with -fmad=false), IEEE division and square root, so the decisions equal the reference's
// own source, compiled on a CPU, bit for bit (tests/test_reabsorb.py).
#include "pr_kernels.h"

#include <cub/device/device_radix_sort.cuh>

namespace pr {

namespace {

struct V3 { float x, y, z; };

// Eigen Vector3f reductions: e0 + (e1 + e2)
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }

// Eigen >= 3.3 normalize()
__device__ __forceinline__ V3 normalize3(V3 a) {
  const float z = dot3(a, a);
  if (z > 0.0f) {
    const float n = sqrtf(z);
    a.x = a.x / n;
    a.y = a.y / n;
    a.z = a.z / n;
  }
  return a;
}

__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  V3 r;
  r.x = a.y * b.z - a.z * b.y;
  r.y = a.z * b.x - a.x * b.z;
  r.z = a.x * b.y - a.y * b.x;
  return r;
}

// distP2P (:203-207): a = dx*dx + dy*dy + dz*dz, pow(a, 0.5f)
__device__ __forceinline__ float dist_sq(V3 p, V3 q) {
  return (p.x - q.x) * (p.x - q.x) + (p.y - q.y) * (p.y - q.y) + (p.z - q.z) * (p.z - q.z);
}
__device__ __forceinline__ float dist_p2p(V3 p, V3 q) { return sqrtf(dist_sq(p, q)); }

// Conservative bound for the "is the intersection inside the segment" test |d1 + d2 - len| < 0.001: once a squared
// end-point distance exceeds (len + 0.01)^2 * (1 + 1e-5), d1 + d2 - len is > 0.009 in FP32 whatever the other distance
// is (sqrt and the two additions are monotonic; the inflation covers their four roundings), so the exact test is
// false and its square roots need not be taken.  NaN bounds or distances compare false and take the exact path.
__device__ __forceinline__ float reject_bound(float len) {
  const float t = len + 0.01f;
  return t * t * 1.00001f;
}

// projPoint2Plane (:1442-1448): float lambda, the subtraction in double
__device__ __forceinline__ V3 proj_point(V3 p, float a, float b, float c, float d) {
  const float inner = a * p.x + b * p.y + c * p.z + d;
  const float lambda = (float)(2.0 * (double)inner);
  const double half = (double)lambda / 2.0;
  V3 r;
  r.x = (float)((double)p.x - half * (double)a);
  r.y = (float)((double)p.y - half * (double)b);
  r.z = (float)((double)p.z - half * (double)c);
  return r;
}

// isBothLineSegsIntersect (:1957-2016) with the edge-only, ray-only and (point, edge)-only quantities hoisted by the
// callers (nab, dist_ab per border edge; ncd, dist_cd per ray; pa_pc and nab.pa_pc per point and edge): the hoisted
// values are computed by the same operations, so every intermediate equals the reference's.
__device__ __forceinline__ bool segs_intersect(V3 pa, V3 pb, V3 nab, float dist_ab, float bound_ab, V3 pc, V3 pd, V3 ncd,
                                               float dist_cd, float bound_cd, V3 pa_pc, float nab_papc) {
  float lambda1, lambda2;
  const float nn = dot3(nab, ncd);
  const float ncd_papc = dot3(ncd, pa_pc);
  if (fabsf(nn) <= 0.001f) {
    lambda1 = nab_papc;
    lambda2 = -1.0f * ncd_papc;
  } else if (nn >= 0.9999f) {
    return false;
  } else {
    const float c1 = 1.0f - nn * nn;
    const float c2 = nab_papc * nn - ncd_papc;
    lambda2 = c2 / c1;
    lambda1 = (lambda2 + ncd_papc) / nn;
  }
  V3 p1, p2, pi;
  p1.x = pa.x + lambda1 * nab.x;
  p1.y = pa.y + lambda1 * nab.y;
  p1.z = pa.z + lambda1 * nab.z;
  p2.x = pc.x + lambda2 * ncd.x;
  p2.y = pc.y + lambda2 * ncd.y;
  p2.z = pc.z + lambda2 * ncd.z;
  pi.x = (p1.x + p2.x) / 2.0f;
  pi.y = (p1.y + p2.y) / 2.0f;
  pi.z = (p1.z + p2.z) / 2.0f;
  const float a_pa = dist_sq(pi, pa), a_pb = dist_sq(pi, pb);
  if (a_pa > bound_ab || a_pb > bound_ab) return false;
  const float a_pc = dist_sq(pi, pc), a_pd = dist_sq(pi, pd);
  if (a_pc > bound_cd || a_pd > bound_cd) return false;
  const float dist_pa = sqrtf(a_pa), dist_pb = sqrtf(a_pb);
  const float dist_pc = sqrtf(a_pc), dist_pd = sqrtf(a_pd);
  return fabsf(dist_pa + dist_pb - dist_ab) < 0.001f && fabsf(dist_pc + dist_pd - dist_cd) < 0.001f;
}

// ---- per-call constants: border edges and ray directions ----------------------------------------------------
// edges: 3 float4 per border vertex e: (pa, dist_ab), (pb, reject bound), (nab, 0); rays: float4 per (plane, ray).
__global__ void __launch_bounds__(128) reabsorb_prepare_kernel(const float4* __restrict__ border, const ReabsorbPlane* __restrict__ planes,
                                                               const int32_t* __restrict__ ray_edges, float4* __restrict__ edges,
                                                               float4* __restrict__ rays) {
  const ReabsorbPlane pl = planes[blockIdx.x];
  const int nb = pl.border_size;
  for (int e = threadIdx.x; e < nb; e += blockDim.x) {
    const float4 a4 = border[pl.border_begin + e];
    const float4 b4 = border[pl.border_begin + (e == nb - 1 ? 0 : e + 1)];
    const V3 pa = {a4.x, a4.y, a4.z}, pb = {b4.x, b4.y, b4.z};
    V3 nab = {pb.x - pa.x, pb.y - pa.y, pb.z - pa.z};
    nab = normalize3(nab);
    const size_t o = 3 * (size_t)(pl.border_begin + e);
    const float dist_ab = dist_p2p(pa, pb);
    edges[o] = make_float4(pa.x, pa.y, pa.z, dist_ab);
    edges[o + 1] = make_float4(pb.x, pb.y, pb.z, reject_bound(dist_ab));
    edges[o + 2] = make_float4(nab.x, nab.y, nab.z, 0.f);
  }
  if (threadIdx.x < 10) {
    // isPointInPoly (:1917-1929): the ray direction is the drawn edge's direction crossed with the plane normal
    const int index = ray_edges[blockIdx.x * 10 + threadIdx.x];
    const float4 s4 = border[pl.border_begin + index];
    const float4 e4 = border[pl.border_begin + (index == nb - 1 ? 0 : index + 1)];
    V3 line_dir = {e4.x - s4.x, e4.y - s4.y, e4.z - s4.z};
    line_dir = normalize3(line_dir);
    const V3 norm = {pl.a, pl.b, pl.c};
    V3 line_dir_p = cross3(line_dir, norm);
    line_dir_p = normalize3(line_dir_p);
    rays[blockIdx.x * 10 + threadIdx.x] = make_float4(line_dir_p.x, line_dir_p.y, line_dir_p.z, 0.f);
  }
}

// ---- R1: distance filter --------------------------------------------------------------------------------------
constexpr int kFilterThreads = 256;
constexpr int kFilterMaxPlanesSmem = 512;

__global__ void __launch_bounds__(kFilterThreads) reabsorb_filter_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                                         const float* __restrict__ Z, size_t n,
                                                                         const ReabsorbPlane* __restrict__ planes, int n_planes, float t,
                                                                         unsigned long long* __restrict__ counter, uint2* __restrict__ cand,
                                                                         unsigned long long cap) {
  __shared__ float4 s_pl[kFilterMaxPlanesSmem];
  for (int j = threadIdx.x; j < n_planes && j < kFilterMaxPlanesSmem; j += blockDim.x)
    s_pl[j] = make_float4(planes[j].a, planes[j].b, planes[j].c, planes[j].d);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t nvec = (n + 3) / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // every lane of a warp runs the same number of iterations (the warp scan below needs all of them)
  const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x - lane;
  for (size_t v0 = first; v0 < nvec; v0 += stride) {
    const size_t v = v0 + lane;
    float xs[4], ys[4], zs[4];
    if (v < nvec) {  // planes are NaN-padded to a multiple of 1024 points: the 128-bit loads stay inside
      const float4 x4 = __ldg(reinterpret_cast<const float4*>(X) + v);
      const float4 y4 = __ldg(reinterpret_cast<const float4*>(Y) + v);
      const float4 z4 = __ldg(reinterpret_cast<const float4*>(Z) + v);
      xs[0] = x4.x; xs[1] = x4.y; xs[2] = x4.z; xs[3] = x4.w;
      ys[0] = y4.x; ys[1] = y4.y; ys[2] = y4.z; ys[3] = y4.w;
      zs[0] = z4.x; zs[1] = z4.y; zs[2] = z4.z; zs[3] = z4.w;
    }
    for (int j0 = 0; j0 < n_planes; j0 += 32) {
      unsigned m[4] = {0u, 0u, 0u, 0u};
      const int jn = min(32, n_planes - j0);
      if (v < nvec) {
        for (int jj = 0; jj < jn; ++jj) {
          const int j = j0 + jj;
          float4 pl;
          if (j < kFilterMaxPlanesSmem) pl = s_pl[j];
          else pl = make_float4(planes[j].a, planes[j].b, planes[j].c, planes[j].d);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const V3 p = {xs[e], ys[e], zs[e]};
            const V3 q = proj_point(p, pl.x, pl.y, pl.z, pl.w);
            const float dist = dist_p2p(p, q);
            // "if (dist > T) return false": NaN distances pass here and fail every segment test later
            if (!(dist > t) && (4 * v + e) < n) m[e] |= 1u << jj;
          }
        }
      }
      const int mine = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
      }
      const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      if (total == 0) continue;
      unsigned long long base = 0;
      if (lane == 31) base = atomicAdd(counter, (unsigned long long)total);
      base = __shfl_sync(0xFFFFFFFFu, base, 31);
      unsigned long long at = base + (unsigned long long)(incl - mine);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        unsigned bits = m[e];
        while (bits) {
          const int jj = __ffs(bits) - 1;
          bits &= bits - 1;
          if (at < cap) cand[at] = make_uint2((unsigned)(4 * v + e), (unsigned)(j0 + jj));
          ++at;
        }
      }
    }
  }
}

// ---- R2: polygon containment ----------------------------------------------------------------------------------
constexpr int kPolyWarps = 8;

__global__ void __launch_bounds__(kPolyWarps * 32) reabsorb_poly_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                                        const float* __restrict__ Z, const uint2* __restrict__ cand,
                                                                        unsigned long long n_cand, const ReabsorbPlane* __restrict__ planes,
                                                                        const float4* __restrict__ edges, const float4* __restrict__ rays,
                                                                        uint32_t* __restrict__ claimed, unsigned long long* __restrict__ keys,
                                                                        unsigned long long* __restrict__ n_absorbed,
                                                                        int32_t* __restrict__ plane_counts) {
  __shared__ __align__(16) float s_ray[kPolyWarps][10][8];  // far.xyz, ncd.xyz, dist_cd, reject bound
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long n_warps = (unsigned long long)gridDim.x * kPolyWarps;
  for (unsigned long long c = (unsigned long long)blockIdx.x * kPolyWarps + warp; c < n_cand; c += n_warps) {
    const uint2 cd = cand[c];
    const ReabsorbPlane pl = planes[cd.y];
    const V3 p = {X[cd.x], Y[cd.x], Z[cd.x]};
    const V3 pc = proj_point(p, pl.a, pl.b, pl.c, pl.d);  // p_base
    __syncwarp();
    if (lane < 10) {
      const float4 dir = rays[cd.y * 10 + lane];
      const float lambda = 10000.0f;
      V3 far;
      far.x = pc.x + lambda * dir.x;
      far.y = pc.y + lambda * dir.y;
      far.z = pc.z + lambda * dir.z;
      V3 ncd = {far.x - pc.x, far.y - pc.y, far.z - pc.z};
      ncd = normalize3(ncd);
      float* r = s_ray[warp][lane];
      r[0] = far.x; r[1] = far.y; r[2] = far.z;
      r[3] = ncd.x; r[4] = ncd.y; r[5] = ncd.z;
      r[6] = dist_p2p(pc, far);
      r[7] = reject_bound(r[6]);
    }
    __syncwarp();
    const int nb = pl.border_size;
    unsigned parity = 0u;
    // lanes over border edges, the ten rays in the inner loop (the edge and pa_pc terms are shared by them)
    for (int e = lane; e < nb; e += 32) {
      const float4* eg = edges + 3 * (size_t)(pl.border_begin + e);
      const float4 a4 = __ldg(eg), b4 = __ldg(eg + 1), n4 = __ldg(eg + 2);
      const V3 pa = {a4.x, a4.y, a4.z}, pb = {b4.x, b4.y, b4.z}, nab = {n4.x, n4.y, n4.z};
      V3 pa_pc;
      pa_pc.x = pc.x - pa.x;
      pa_pc.y = pc.y - pa.y;
      pa_pc.z = pc.z - pa.z;
      const float nab_papc = dot3(nab, pa_pc);
#pragma unroll 2
      for (int r = 0; r < 10; ++r) {
        const float4 r0 = *reinterpret_cast<const float4*>(&s_ray[warp][r][0]);
        const float4 r1 = *reinterpret_cast<const float4*>(&s_ray[warp][r][4]);
        const V3 pd = {r0.x, r0.y, r0.z}, ncd = {r0.w, r1.x, r1.y};
        if (segs_intersect(pa, pb, nab, a4.w, b4.w, pc, pd, ncd, r1.z, r1.w, pa_pc, nab_papc)) parity ^= 1u << r;
      }
    }
    parity = __reduce_xor_sync(0xFFFFFFFFu, parity);
    if (lane == 0 && __popc(parity) >= 5) {  // count >= count_for_intersect.size() / 2
      claimed[cd.x] = 1u;
      const unsigned long long slot = atomicAdd(n_absorbed, 1ull);
      keys[slot] = ((unsigned long long)cd.y << 32) | (unsigned long long)cd.x;
      atomicAdd(&plane_counts[cd.y], 1);
    }
  }
}

// ---- R3: sorted keys -> index lists ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reabsorb_split_kernel(const unsigned long long* __restrict__ keys, size_t n,
                                                             const int32_t* __restrict__ orig, int32_t* __restrict__ out_cur,
                                                             int32_t* __restrict__ out_orig) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t idx = (int32_t)(keys[i] & 0xFFFFFFFFull);
    out_cur[i] = idx;
    out_orig[i] = orig ? orig[idx] : idx;
  }
}

}  // namespace

void launch_reabsorb_prepare(const float4* border, const ReabsorbPlane* planes, int n_planes, const int32_t* ray_edges,
                             float4* edges, float4* rays, cudaStream_t s) {
  if (n_planes <= 0) return;
  reabsorb_prepare_kernel<<<n_planes, 128, 0, s>>>(border, planes, ray_edges, edges, rays);
}

void launch_reabsorb_filter(CloudView cloud, size_t n, const ReabsorbPlane* planes, int n_planes, float t,
                            unsigned long long* counter, uint2* cand, unsigned long long cap, int num_sms, cudaStream_t s) {
  if (n == 0 || n_planes <= 0) return;
  const size_t nvec = (n + 3) / 4;
  size_t blocks = (nvec + kFilterThreads - 1) / kFilterThreads;
  if (blocks > (size_t)num_sms * 8) blocks = (size_t)num_sms * 8;
  reabsorb_filter_kernel<<<(unsigned)blocks, kFilterThreads, 0, s>>>(cloud.x, cloud.y, cloud.z, n, planes, n_planes, t, counter, cand, cap);
}

void launch_reabsorb_poly(CloudView cloud, const uint2* cand, unsigned long long n_cand, const ReabsorbPlane* planes,
                          const float4* edges, const float4* rays, uint32_t* claimed, unsigned long long* keys,
                          unsigned long long* n_absorbed, int32_t* plane_counts, int num_sms, cudaStream_t s) {
  if (n_cand == 0) return;
  unsigned long long blocks = (n_cand + kPolyWarps - 1) / kPolyWarps;
  if (blocks > (unsigned long long)num_sms * 8) blocks = (unsigned long long)num_sms * 8;
  reabsorb_poly_kernel<<<(unsigned)blocks, kPolyWarps * 32, 0, s>>>(cloud.x, cloud.y, cloud.z, cand, n_cand, planes, edges, rays, claimed,
                                                                   keys, n_absorbed, plane_counts);
}

size_t reabsorb_sort_temp_bytes(size_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)n);
  return bytes;
}

void launch_reabsorb_lists(const unsigned long long* keys_in, unsigned long long* keys_sorted, size_t n, int key_bits, void* temp,
                           size_t temp_bytes, const int32_t* orig, int32_t* out_cur, int32_t* out_orig, cudaStream_t s) {
  if (n == 0) return;
  cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys_in, keys_sorted, (int)n, 0, key_bits, s);
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  reabsorb_split_kernel<<<(unsigned)blocks, 256, 0, s>>>(keys_sorted, n, orig, out_cur, out_orig);
}

}  // namespace pr
