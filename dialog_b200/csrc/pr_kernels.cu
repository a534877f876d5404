// pr_kernels.cu — sm_100a kernels of the plane-RANSAC backend (see DESIGN.md for the rooflines).
//
//   K0 stage_kernel     AoS pcl::PointXYZ -> x[] y[] z[] planes, NaN padding, bounding box
//   K1 gather/models    sample triples -> plane hypotheses, PCL op order, no contraction
//   K2 score_kernel     N x K threshold tests; FP32-FMA bound; TMA bulk copies into a smem ring
//   K3 refit_kernel     inlier predicate + exact integer moments; HBM bound
//   K5 compact_kernel   stable partition (decoupled look-back scan); HBM bound
//
// Reference arithmetic restated here: the point-to-plane threshold test of
// Dialog/PlaneDetect.h:1442-1448,1902,2019-2023 (PCL countWithinDistance), the order-preserving
// rebuild of source_cloud at Dialog/PlaneDetect.h:1560-1566 (PCL ExtractIndices), and the
// covariance accumulation behind pcl::computePointNormal at Dialog/PlaneDetect.h:1485.
//
// Built with -fmad=false: every multiply-add that is fused is written as an explicit fma intrinsic,
// so the FP32 results are a function of the source text only.
#include "pr_kernels.h"

#include <math_constants.h>

#include "pr_chain_dev.cuh"

#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

namespace pr {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// coeff . (x, y, z, 1) in the two documented orders (include/plane_ransac.h PR_DOT_*).
template <int DOT>
__device__ __forceinline__ float plane_dot(float a, float b, float c, float d, float x, float y, float z) {
  if (DOT == 1) return __fmaf_rn(a, x, __fmaf_rn(b, y, __fmaf_rn(c, z, d)));
  return __fadd_rn(__fadd_rn(__fmul_rn(a, x), __fmul_rn(c, z)), __fadd_rn(__fmul_rn(b, y), d));
}

template <int DOT>
__device__ __forceinline__ float2 plane_dot2(float a, float b, float c, float d, float2 x, float2 y, float2 z) {
  if (DOT == 1) {
    const float2 A = make_float2(a, a), B = make_float2(b, b), C = make_float2(c, c), D = make_float2(d, d);
    return __ffma2_rn(A, x, __ffma2_rn(B, y, __ffma2_rn(C, z, D)));
  }
  // Separately rounded products and sums must stay scalar: ptxas 12.9 contracts mul.rn.f32x2 +
  // add.rn.f32x2 into FFMA2 even though both carry an explicit rounding mode (seen in SASS and as
  // +-1 count differences in the parity tests); scalar FMUL/FADD with .rn are never fused.
  return make_float2(plane_dot<0>(a, b, c, d, x.x, y.x, z.x), plane_dot<0>(a, b, c, d, x.y, y.y, z.y));
}

__device__ __forceinline__ uint32_t float_order_key(float f) {
  uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// ------------------------------------------------------------------------------------------------
// K0: staging
// ------------------------------------------------------------------------------------------------
__global__ void bbox_init_kernel(uint32_t* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0xFFFFFFFFu;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) stage_kernel(const float4* __restrict__ aos, size_t n, float* __restrict__ x,
                                                    float* __restrict__ y, float* __restrict__ z, size_t cap,
                                                    uint32_t* __restrict__ bbox) {
  uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
    if (i < n) {
      float4 p = __ldg(&aos[i]);
      x[i] = p.x;
      y[i] = p.y;
      z[i] = p.z;
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const float c[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          uint32_t k = float_order_key(c[a]);
          lo[a] = min(lo[a], k);
          hi[a] = max(hi[a], k);
        }
      }
    } else {
      x[i] = CUDART_NAN_F;
      y[i] = CUDART_NAN_F;
      z[i] = CUDART_NAN_F;
    }
  }
  // one atomic per block and bound: atomics on the same six words serialise in L2 (a few ns each), which at one per
  // warp cost more than moving the chunk itself
  __shared__ uint32_t s_key[6][8];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    uint32_t l = __reduce_min_sync(0xFFFFFFFFu, lo[a]);
    uint32_t h = __reduce_max_sync(0xFFFFFFFFu, hi[a]);
    if ((threadIdx.x & 31) == 0) {
      s_key[a][threadIdx.x >> 5] = l;
      s_key[3 + a][threadIdx.x >> 5] = h;
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const bool is_min = threadIdx.x < 3;
    uint32_t k = s_key[threadIdx.x][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) k = is_min ? min(k, s_key[threadIdx.x][w]) : max(k, s_key[threadIdx.x][w]);
    if (is_min) {
      if (k != 0xFFFFFFFFu) atomicMin(&bbox[threadIdx.x], k);
    } else if (k != 0u) {
      atomicMax(&bbox[threadIdx.x], k);
    }
  }
}

__global__ void __launch_bounds__(256) unstage_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      const float* __restrict__ z, size_t n, float4* __restrict__ aos) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    aos[i] = make_float4(x[i], y[i], z[i], 1.0f);
}

// Batch of equal-sized clouds: cloud c occupies [c * stride, c * stride + n_per) of each plane, NaN up to
// the next cloud; one bounding box per cloud.
__global__ void bbox_init_batch_kernel(uint32_t* bbox, int n_clouds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_clouds * 6) bbox[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;
}

__global__ void __launch_bounds__(256) stage_batch_kernel(const float4* __restrict__ aos, size_t n_per, size_t stride,
                                                          float* __restrict__ x, float* __restrict__ y,
                                                          float* __restrict__ z, size_t tail, uint32_t* __restrict__ bbox) {
  const size_t c = blockIdx.y;
  const bool last = c + 1 == gridDim.y;
  const size_t span = stride + (last ? tail : 0);
  uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < span; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = c * stride + i;
    if (i < n_per) {
      float4 p = __ldg(&aos[c * n_per + i]);
      x[o] = p.x;
      y[o] = p.y;
      z[o] = p.z;
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const float v[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          uint32_t k = float_order_key(v[a]);
          lo[a] = min(lo[a], k);
          hi[a] = max(hi[a], k);
        }
      }
    } else {
      x[o] = CUDART_NAN_F;
      y[o] = CUDART_NAN_F;
      z[o] = CUDART_NAN_F;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    uint32_t l = __reduce_min_sync(0xFFFFFFFFu, lo[a]);
    uint32_t h = __reduce_max_sync(0xFFFFFFFFu, hi[a]);
    if ((threadIdx.x & 31) == 0) {
      if (l != 0xFFFFFFFFu) atomicMin(&bbox[c * 6 + a], l);
      if (h != 0u) atomicMax(&bbox[c * 6 + 3 + a], h);
    }
  }
}

void launch_stage_batch(const float4* aos, size_t n_clouds, size_t n_per, size_t stride, CloudView dst, uint32_t* bbox,
                        cudaStream_t s) {
  bbox_init_batch_kernel<<<(unsigned)((n_clouds * 6 + 255) / 256), 256, 0, s>>>(bbox, (int)n_clouds);
  unsigned bx = (unsigned)((stride + kTilePoints + 255) / 256);
  if (bx > 64) bx = 64;
  dim3 grid(bx, (unsigned)n_clouds);
  stage_batch_kernel<<<grid, 256, 0, s>>>(aos, n_per, stride, dst.x, dst.y, dst.z, dst.cap - n_clouds * stride, bbox);
}

void launch_bbox_init(uint32_t* bbox, cudaStream_t s) { bbox_init_kernel<<<1, 32, 0, s>>>(bbox); }

// Exact centroid (preProcess's translation, Dialog/PlaneDetect.h:458-481, made order-independent):
// sums[0..3) = sum of rint((p - lo) * 2^s) over the finite points, sums[3] = their number.
__global__ void __launch_bounds__(256) centroid_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                       const float* __restrict__ Z, size_t n, double lox, double loy,
                                                       double loz, double scale, long long* __restrict__ sums) {
  long long acc[4] = {0, 0, 0, 0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float x = X[i], y = Y[i], z = Z[i];
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      acc[0] += __double2ll_rn(((double)x - lox) * scale);
      acc[1] += __double2ll_rn(((double)y - loy) * scale);
      acc[2] += __double2ll_rn(((double)z - loz) * scale);
      acc[3] += 1;
    }
  }
  __shared__ long long s_part[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    long long v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    long long v = 0;
    for (int w = 0; w < 8; ++w) v += s_part[w][threadIdx.x];
    if (v) atomicAdd(reinterpret_cast<unsigned long long*>(&sums[threadIdx.x]), (unsigned long long)v);
  }
}

__global__ void __launch_bounds__(256) translate_kernel(float* __restrict__ X, float* __restrict__ Y, float* __restrict__ Z,
                                                        size_t n, float cx, float cy, float cz) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    X[i] = __fsub_rn(X[i], cx);  // points[i].x -= p.x, one float subtraction per coordinate
    Y[i] = __fsub_rn(Y[i], cy);
    Z[i] = __fsub_rn(Z[i], cz);
  }
}

void launch_centroid(CloudView c, size_t n, const double lo[3], int scale_exp, long long* sums, int num_sms, cudaStream_t s) {
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)num_sms * 8) blocks = (size_t)num_sms * 8;
  centroid_kernel<<<(unsigned)blocks, 256, 0, s>>>(c.x, c.y, c.z, n, lo[0], lo[1], lo[2], ldexp(1.0, scale_exp), sums);
}

void launch_translate(CloudView c, size_t n, const float centroid[3], int num_sms, cudaStream_t s) {
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)num_sms * 8) blocks = (size_t)num_sms * 8;
  translate_kernel<<<(unsigned)blocks, 256, 0, s>>>(c.x, c.y, c.z, n, centroid[0], centroid[1], centroid[2]);
}

void launch_stage(const float4* aos, size_t n, CloudView dst, uint32_t* bbox, cudaStream_t s) {
  size_t blocks = (dst.cap + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;  // one resident wave
  stage_kernel<<<(unsigned)blocks, 256, 0, s>>>(aos, n, dst.x, dst.y, dst.z, dst.cap, bbox);
}

// Plane::points_set for one plane (gather by inlier index), optionally projected onto the plane with the
// reference's projPoint2Plane arithmetic (Dialog/PlaneDetect.h:1442-1448): lambda = 2.0 * (a*x + b*y + c*z + d)
// in float, p' = p - lambda / 2.0 * n evaluated in double and rounded to float.
__global__ void __launch_bounds__(256) plane_points_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           const float* __restrict__ z, const int32_t* __restrict__ idx,
                                                           size_t n, Plane4 pl, bool project, float4* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t j = idx[i];
    float px = x[j], py = y[j], pz = z[j];
    if (project) {
      const float inner = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pl.a, px), __fmul_rn(pl.b, py)), __fmul_rn(pl.c, pz)), pl.d);
      const float lambda = (float)(2.0 * (double)inner);
      const double half = (double)lambda / 2.0;
      px = (float)__dsub_rn((double)px, __dmul_rn(half, (double)pl.a));
      py = (float)__dsub_rn((double)py, __dmul_rn(half, (double)pl.b));
      pz = (float)__dsub_rn((double)pz, __dmul_rn(half, (double)pl.c));
    }
    out[i] = make_float4(px, py, pz, 1.0f);
  }
}

void launch_plane_points(CloudView cloud, const int32_t* idx, size_t n, Plane4 pl, bool project, float4* out, cudaStream_t s) {
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  plane_points_kernel<<<(unsigned)blocks, 256, 0, s>>>(cloud.x, cloud.y, cloud.z, idx, n, pl, project, out);
}

// out[i] = map ? map[idx[i]] : idx[i]  (composition of index maps when a peeled cloud becomes the staged cloud)
__global__ void __launch_bounds__(256) compose_map_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ map, size_t n,
                                                          int32_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t j = idx[i];
    out[i] = map ? map[j] : j;
  }
}

void launch_compose_map(const int32_t* idx, const int32_t* map, size_t n, int32_t* out, cudaStream_t s) {
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  compose_map_kernel<<<(unsigned)blocks, 256, 0, s>>>(idx, map, n, out);
}

void launch_unstage(CloudView src, size_t n, float4* aos, cudaStream_t s) {
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  unstage_kernel<<<(unsigned)blocks, 256, 0, s>>>(src.x, src.y, src.z, n, aos);
}

// ------------------------------------------------------------------------------------------------
// K1: hypotheses from sample triples
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_samples_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ z, long long first, size_t n,
                                                             const int32_t* __restrict__ triples, int n_samples,
                                                             int4* __restrict__ out, int n_clouds, size_t cloud_stride,
                                                             bool per_cloud_triples, const RoundState* __restrict__ st) {
  if (st != nullptr) {  // peel loop without the host: this round's shard extent lives on the device
    if (st->stop) return;
    first = st->first;
    n = (size_t)st->n_local;
  }
  const long long total = (long long)n_samples * n_clouds;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (long long)gridDim.x * blockDim.x) {
    int c = (int)(s / n_samples);
    int j = (int)(s - (long long)c * n_samples);
    long long local = (long long)triples[per_cloud_triples ? s : j] - first;
    int4 v = make_int4(0, 0, 0, 0);
    if (local >= 0 && local < (long long)n) {
      size_t o = (size_t)c * cloud_stride + (size_t)local;
      v = make_int4(__float_as_int(x[o]), __float_as_int(y[o]), __float_as_int(z[o]), 0x3F800000);
    }
    out[s] = v;
  }
}

// model_from_sample (PCL isSampleGood + computeModelCoefficients) lives in pr_chain_dev.cuh: the exchange kernels use it too.
__global__ void __launch_bounds__(128) models_kernel(const int4* __restrict__ sample_pts, int n_models,
                                                     float4* __restrict__ hyps, int32_t* __restrict__ good) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_models) return;
  float4 h;
  const bool ok = model_from_sample(sample_pts[3 * k], sample_pts[3 * k + 1], sample_pts[3 * k + 2], &h);
  hyps[k] = h;
  good[k] = ok ? 1 : 0;
}

// K1a + K1b in one launch for a cloud on one GPU driven by the device-resident round state: every thread gathers its
// three sample points (kept in sample_pts: the refit takes its pivot from there) and forms the model.
__global__ void __launch_bounds__(128) gather_models_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ z, const int32_t* __restrict__ triples,
                                                            int n_models, int4* __restrict__ sample_pts, float4* __restrict__ hyps,
                                                            int32_t* __restrict__ good, const RoundState* __restrict__ st) {
  pdl_wait();
  if (st->stop) return;
  chain_stamp(st, kStampModels);
  const long long n = st->n_local;
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_models) return;
  int4 q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const long long j = (long long)triples[3 * k + i];
    q[i] = make_int4(0, 0, 0, 0);
    if (j >= 0 && j < n) q[i] = make_int4(__float_as_int(x[j]), __float_as_int(y[j]), __float_as_int(z[j]), 0x3F800000);
    sample_pts[3 * k + i] = q[i];
  }
  float4 h;
  const bool ok = model_from_sample(q[0], q[1], q[2], &h);
  hyps[k] = h;
  good[k] = ok ? 1 : 0;
}

void launch_gather_models(CloudView cloud, const int32_t* triples, int n_models, int4* sample_pts, float4* hyps, int32_t* good,
                          const RoundState* st, cudaStream_t s) {
  if (n_models <= 0) return;
  launch_chained(gather_models_kernel, dim3((n_models + 127) / 128), dim3(128), 0, s, cloud.x, cloud.y, cloud.z, triples, n_models, sample_pts,
                 hyps, good, st);
}

void launch_gather_samples(CloudView cloud, long long first, size_t n, const int32_t* triples, int n_samples,
                           int4* sample_pts, int n_clouds, size_t cloud_stride, cudaStream_t s, bool per_cloud_triples,
                           const RoundState* st) {
  long long total = (long long)n_samples * n_clouds;
  if (total <= 0) return;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  gather_samples_kernel<<<(unsigned)blocks, 256, 0, s>>>(cloud.x, cloud.y, cloud.z, first, n, triples, n_samples,
                                                         sample_pts, n_clouds, cloud_stride, per_cloud_triples, st);  // first of its sequence: plain
}

void launch_models(const int4* sample_pts, int n_models_total, float4* hyps, int32_t* good, cudaStream_t s) {
  if (n_models_total <= 0) return;
  models_kernel<<<(n_models_total + 127) / 128, 128, 0, s>>>(sample_pts, n_models_total, hyps, good);
}

// ------------------------------------------------------------------------------------------------
// K2: hypothesis scoring
//
// Layout: hypothesis-per-lane, point broadcast.  Each lane keeps H hypotheses (a,b,c,d) in registers
// for the whole kernel; a warp walks a tile of points held in shared memory, reading four points per
// step with three broadcast LDS.128 (x[4], y[4], z[4]); every lane evaluates its H hypotheses on the
// four points with packed FP32 FMAs (FFMA2: two points per instruction, the hypothesis operand
// broadcast), tests |r| < t with FSET.BF and adds the 1.0f/0.0f bit patterns two at a time with one
// IADD3.  Counts stay lane-private until the end: no shuffles, no shared-memory atomics.
//
// The 1.0f pattern is 0x3F800000 = 127 * 2^23, so after c increments the accumulator holds
// (127 c mod 512) << 23; c < 512 is recovered as ((acc >> 23) * 383) & 511 (127 * 383 = 1 mod 512) and
// folded into a plain counter every <= 64 steps (<= 256 increments).
//
// Tiles arrive through a kScoreStages-deep ring of TMA bulk copies completing on "full" mbarriers.
// There is no producer warp: the last warp to finish a stage (shared-memory ticket) re-arms its
// barrier and issues the copies for the tile kScoreStages ahead, so nothing ever spins on an empty
// slot (a spinning producer lane cost ~17 % of the issue slots of its SM sub-partition in the first
// ncu capture).  The work list is (cloud, tile) items so that a batch of small clouds runs in one launch.
// ------------------------------------------------------------------------------------------------
constexpr int kScoreWarps = 8;
constexpr int kScoreThreads = kScoreWarps * 32;
constexpr int kScoreStages = 3;
constexpr int kScoreStepsPerFlush = 32;             // 128 points: <= 128 increments per counter between folds
constexpr int kScoreStageFloats = 3 * kTilePoints + 4;  // + 16 bytes so the last prefetch stays inside the stage

// One tile (`len` points at element offset `off` of the planes) -> three bulk copies into stage `s`,
// completion on s_full[s].
__device__ __forceinline__ void score_issue_tile(const float* X, const float* Y, const float* Z, size_t off, int len,
                                                 float* stage, uint64_t* full) {
  mbar_expect_tx(full, 3u * (unsigned)len * sizeof(float));
  tma_bulk_g2s(stage, X + off, (unsigned)len * sizeof(float), full);
  tma_bulk_g2s(stage + kTilePoints, Y + off, (unsigned)len * sizeof(float), full);
  tma_bulk_g2s(stage + 2 * kTilePoints, Z + off, (unsigned)len * sizeof(float), full);
}

// Work of one CTA.  Item mode (pts_per_cta == 0, batches of small clouds): items are (cloud, 1024-point tile)
// pairs, items_per_cta of them per CTA.  Range mode (one cloud): the CTA owns the points
// [blockIdx.x * pts_per_cta, +pts_per_cta) cut into 1024-point tiles with a shorter last one, so that the
// cloud divides evenly over the CTA slots of the launch whatever its size (a whole-tile split loses up to
// 1/(tiles per CTA) of the machine on the rounding, 6-8 % on a 2M-point round).
struct ScoreSpan {
  int n_items;
  int item_begin;       // item mode
  long long pt_begin;   // range mode
  int pts;              // range mode: points of this CTA (multiple of the warp-split unit)
};

__device__ __forceinline__ void score_tile_of(const ScoreSpan& w, int it, int tiles_per_cloud, size_t cloud_stride, int pts_per_cta,
                                              int* cloud, size_t* off, int* len) {
  if (pts_per_cta == 0) {
    const int item = w.item_begin + it;
    const int c = item / tiles_per_cloud, tl = item - c * tiles_per_cloud;
    *cloud = c;
    *off = (size_t)c * cloud_stride + (size_t)tl * kTilePoints;
    *len = kTilePoints;
  } else {
    *cloud = 0;
    *off = (size_t)w.pt_begin + (size_t)it * kTilePoints;
    *len = min(kTilePoints, w.pts - it * kTilePoints);
  }
}

template <int H, int DOT, int UNR>
__global__ void __launch_bounds__(kScoreThreads, H <= 4 ? 4 : 2)
    score_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Z,
                 size_t cloud_stride, int tiles_per_cloud, int total_items, int items_per_cta, int pts_per_cta,
                 long long n_padded, const float4* __restrict__ hyps, int K, int k_begin, int k_end, float t,
                 int32_t* __restrict__ counts, int warps_h, const RoundState* __restrict__ st) {
  // hypotheses [k_begin, k_end) of the K per cloud are scored by this launch
  pdl_wait();  // returns at once unless launched as a programmatic dependent (queued rounds, batch path)
  if (st != nullptr) {
    // peel loop without the host (range mode): the even split of launch_score_h, from the cloud size on the device
    if (st->stop) return;
    if (k_begin == 0) chain_stamp(st, kStampScore);
    const long long unit = 128ll * (kScoreWarps / warps_h);
    n_padded = (st->n_local + unit - 1) / unit * unit;
    long long per = (n_padded + gridDim.x - 1) / gridDim.x;
    per = (per + unit - 1) / unit * unit;
    pts_per_cta = (int)per;
    if (n_padded == 0) return;
  }
  __shared__ __align__(128) float s_pts[kScoreStages][kScoreStageFloats];
  __shared__ __align__(8) uint64_t s_full[kScoreStages];
  __shared__ int s_done[kScoreStages];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ScoreSpan w;
  if (pts_per_cta == 0) {
    w.item_begin = blockIdx.x * items_per_cta;
    w.n_items = min(total_items, w.item_begin + items_per_cta) - w.item_begin;
    w.pt_begin = 0;
    w.pts = 0;
  } else {
    w.item_begin = 0;
    w.pt_begin = (long long)blockIdx.x * pts_per_cta;
    w.pts = (int)min((long long)pts_per_cta, n_padded - w.pt_begin);
    w.n_items = (w.pts + kTilePoints - 1) / kTilePoints;
  }
  const int n_items = w.n_items;
  if (n_items <= 0) return;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kScoreStages; ++s) {
      mbar_init(&s_full[s], 1);
      s_done[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // prologue: fill the ring
    for (int s = 0; s < kScoreStages && s < n_items; ++s) {
      int c_, len_;
      size_t off_;
      score_tile_of(w, s, tiles_per_cloud, cloud_stride, pts_per_cta, &c_, &off_, &len_);
      score_issue_tile(X, Y, Z, off_, len_, &s_pts[s][0], &s_full[s]);
    }
  }
  __syncthreads();

  const int wh = warp % warps_h, wp = warp / warps_h;
  const int warps_p = kScoreWarps / warps_h;
  const int k_stride = 32 * warps_h;
  const int k0 = k_begin + blockIdx.y * (k_stride * H) + wh * 32 + lane;

  float ha[H], hb[H], hc[H], hd[H];
  int cnt[H];
  unsigned acc[H];
  int cur_cloud = -1;
  // item mode: (cloud, tile) of the item being scored and of the one the ring is refilled with, advanced step by step
  // (a division per tile is ~1 % of a warp's work on a 128-point share)
  int it_c = 0, it_t = 0, nx_c = 0, nx_t = 0;
  if (pts_per_cta == 0) {
    it_c = w.item_begin / tiles_per_cloud;
    it_t = w.item_begin - it_c * tiles_per_cloud;
    const int nx = w.item_begin + kScoreStages;
    nx_c = nx / tiles_per_cloud;
    nx_t = nx - nx_c * tiles_per_cloud;
  }
  int pending = 0;  // 128-point chunks accumulated in the packed counters since the last fold (at most 3)

  for (int it = 0; it < n_items; ++it) {
    const int s = it % kScoreStages;
    int c, len;
    size_t off;
    if (pts_per_cta == 0) {
      c = it_c;
      off = (size_t)it_c * cloud_stride + (size_t)it_t * kTilePoints;
      len = kTilePoints;
    } else {
      score_tile_of(w, it, tiles_per_cloud, cloud_stride, pts_per_cta, &c, &off, &len);
    }
    const int pts_per_warp = len / warps_p;  // a multiple of 128: len is a multiple of 128 * warps_p
    const int p_begin = wp * pts_per_warp;
    if (c != cur_cloud) {
      if (cur_cloud >= 0) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
          cnt[j] += (int)(((acc[j] >> 23) * 383u) & 511u);
          const int k = k0 + j * k_stride;
          if (k < k_end && cnt[j] != 0) atomicAdd(&counts[(size_t)cur_cloud * K + k], cnt[j]);
        }
        pending = 0;
      }
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const int k = k0 + j * k_stride;
        float4 h = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        if (k < k_end) h = __ldg(&hyps[(size_t)c * K + k]);
        ha[j] = h.x; hb[j] = h.y; hc[j] = h.z; hd[j] = h.w;
        cnt[j] = 0;
        acc[j] = 0u;
      }
      cur_cloud = c;
    }

    mbar_wait(&s_full[s], (it / kScoreStages) & 1);
    const float* sx = &s_pts[s][p_begin];
    const float* sy = sx + kTilePoints;
    const float* sz = sx + 2 * kTilePoints;

    // 128-point chunks (32 steps of 4 points): compile-time trip count; the packed counters hold up to 511 increments,
    // so they are folded into the plain counters every third chunk (384 increments), across tiles, and before the
    // counts leave the registers
    for (int p = 0; p < pts_per_warp; p += 4 * kScoreStepsPerFlush) {
      // software pipeline: the points of step q + 1 are loaded while step q computes (the load issued by
      // the last step of a tile reads 16 bytes past its plane: the next plane, or the stage's tail pad)
      float4 x4 = *reinterpret_cast<const float4*>(sx + p);
      float4 y4 = *reinterpret_cast<const float4*>(sy + p);
      float4 z4 = *reinterpret_cast<const float4*>(sz + p);
#pragma unroll UNR
      for (int q = 0; q < kScoreStepsPerFlush; ++q) {
        const float2 xa = make_float2(x4.x, x4.y), xb = make_float2(x4.z, x4.w);
        const float2 ya = make_float2(y4.x, y4.y), yb = make_float2(y4.z, y4.w);
        const float2 za = make_float2(z4.x, z4.y), zb = make_float2(z4.z, z4.w);
        x4 = *reinterpret_cast<const float4*>(sx + p + 4 * q + 4);
        y4 = *reinterpret_cast<const float4*>(sy + p + 4 * q + 4);
        z4 = *reinterpret_cast<const float4*>(sz + p + 4 * q + 4);
#pragma unroll
        for (int j = 0; j < H; ++j) {
          const float2 ra = plane_dot2<DOT>(ha[j], hb[j], hc[j], hd[j], xa, ya, za);
          const float2 rb = plane_dot2<DOT>(ha[j], hb[j], hc[j], hd[j], xb, yb, zb);
          float f0, f1, f2, f3;
          asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f0) : "f"(fabsf(ra.x)), "f"(t));
          asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f1) : "f"(fabsf(ra.y)), "f"(t));
          asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f2) : "f"(fabsf(rb.x)), "f"(t));
          asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f3) : "f"(fabsf(rb.y)), "f"(t));
          unsigned a = acc[j];
          a = a + __float_as_uint(f0) + __float_as_uint(f1);
          a = a + __float_as_uint(f2) + __float_as_uint(f3);
          acc[j] = a;
        }
      }
      if (++pending == 3) {
        pending = 0;
#pragma unroll
        for (int j = 0; j < H; ++j) {
          cnt[j] += (int)(((acc[j] >> 23) * 383u) & 511u);
          acc[j] = 0u;
        }
      }
    }
    if (pts_per_cta == 0) {
      if (++it_t == tiles_per_cloud) { it_t = 0; ++it_c; }
    }

    // release the stage; the last warp to finish refills it (no thread ever spins on an empty slot)
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const int prev = atomicAdd(&s_done[s], 1);
      if (prev == kScoreWarps - 1) {
        s_done[s] = 0;
        const int next = it + kScoreStages;
        if (next < n_items) {
          __threadfence_block();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          int c_, len_;
          size_t off_;
          if (pts_per_cta == 0) {
            off_ = (size_t)nx_c * cloud_stride + (size_t)nx_t * kTilePoints;
            len_ = kTilePoints;
          } else {
            score_tile_of(w, next, tiles_per_cloud, cloud_stride, pts_per_cta, &c_, &off_, &len_);
          }
          score_issue_tile(X, Y, Z, off_, len_, &s_pts[s][0], &s_full[s]);
        }
      }
    }
    if (pts_per_cta == 0) {
      if (++nx_t == tiles_per_cloud) { nx_t = 0; ++nx_c; }
    }
  }
#pragma unroll
  for (int j = 0; j < H; ++j) {
    cnt[j] += (int)(((acc[j] >> 23) * 383u) & 511u);
    const int k = k0 + j * k_stride;
    if (k < k_end && cnt[j] != 0) atomicAdd(&counts[(size_t)cur_cloud * K + k], cnt[j]);
  }
}

// One launch scoring hypotheses [k_begin, k_end) with H per lane and warps_h warps across hypotheses.
template <int H>
static void launch_score_h(const float* X, const float* Y, const float* Z, size_t n_per_cloud, size_t cloud_stride,
                           int n_clouds, const float4* hyps, int K, int k_begin, int k_end, int warps_h, float t,
                           int dot_order, int32_t* counts, int num_sms, cudaStream_t s, bool range_mode, const RoundState* st, bool chained) {
  const int chunk = 32 * H * warps_h;
  const int n_chunks = (k_end - k_begin + chunk - 1) / chunk;
  int slots = ((H <= 4 ? 4 : 2) * num_sms) / n_chunks;
  if (slots < 1) slots = 1;
  const int tiles_per_cloud = (int)((n_per_cloud + kTilePoints - 1) / kTilePoints);
  int total_items = 0, items_per_cta = 0, pts_per_cta = 0, gx;
  long long n_padded = 0;
  if (range_mode) {
    // even split of the (padded) cloud in units every warp of the CTA can share: 128 points per warp
    const long long unit = 128ll * (kScoreWarps / warps_h);
    n_padded = ((long long)n_per_cloud + unit - 1) / unit * unit;
    long long per = (n_padded + slots - 1) / slots;
    per = (per + unit - 1) / unit * unit;
    pts_per_cta = (int)per;
    gx = (int)((n_padded + per - 1) / per);
    if (st != nullptr) gx = slots;  // the device recomputes the split for the round's actual size
  } else {
    total_items = tiles_per_cloud * n_clouds;
    gx = slots > total_items ? total_items : slots;
    items_per_cta = (total_items + gx - 1) / gx;
    gx = (total_items + items_per_cta - 1) / items_per_cta;
  }
  dim3 grid(gx, n_chunks);
  // loop unroll of the H = 8 / FMA-order kernel: % of FP32 peak at N = 10M, K = 4096 measured on B200: 59.2 (2), 60.3 (8),
  // 61.0 (16), 50.9 (32: the body no longer fits the instruction cache) — profiles/r02_score_unroll.txt
#define PR_SCORE(HH, D, U)                                                                                                    \
  do {                                                                                                                        \
    if (st != nullptr || chained)                                                                                             \
      launch_chained(score_kernel<HH, D, U>, grid, dim3(kScoreThreads), 0, s, X, Y, Z, cloud_stride, tiles_per_cloud, total_items,      \
                     items_per_cta, pts_per_cta, n_padded, hyps, K, k_begin, k_end, t, counts, warps_h, st);                  \
    else                                                                                                                      \
      score_kernel<HH, D, U><<<grid, kScoreThreads, 0, s>>>(X, Y, Z, cloud_stride, tiles_per_cloud, total_items, items_per_cta, \
                                                            pts_per_cta, n_padded, hyps, K, k_begin, k_end, t, counts, warps_h, st); \
  } while (0)
  if (dot_order == 1) {
    if (H == 8) PR_SCORE(8, 1, 16);
    else PR_SCORE(H, 1, 2);
  } else {
    PR_SCORE(H, 0, 2);
  }
#undef PR_SCORE
}

int launch_score(CloudView cloud, size_t n_per_cloud, int n_clouds, size_t cloud_stride, const float4* hyps, int K,
                 float t, int dot_order, int32_t* counts, int num_sms, cudaStream_t s, const RoundState* st, bool chained) {
  if (K <= 0 || n_per_cloud == 0 || n_clouds <= 0) return 0;
  static const int forced_h = [] { const char* e = getenv("PR_SCORE_H"); return e ? atoi(e) : 0; }();  // tuning knob
  static const int geom = [] { const char* e = getenv("PR_SCORE_GEOM"); return e ? atoi(e) : 1; }();   // 0: round-1 whole-tile split
  const bool range_mode = (geom != 0 || st != nullptr) && n_clouds == 1;
  if (st != nullptr && !range_mode) return 0;
  // Every lane slot of a launch costs the same whether or not it holds a hypothesis, so K is cut into
  // launches whose slot count (32 * H * warps_h per chunk, grid.y chunks) matches.  With the even point split
  // any chunk count fills the machine, so the largest chunk width that wastes <= 1/32 of its slots takes all of
  // K in one launch (3072 = 3 x 1024); otherwise full 2048-wide chunks first, then powers of two, rounding the
  // last pieces up when little is wasted.
  int launches = 0;
  int k = 0;
  while (k < K) {
    const int rest = K - k;
    int H = 0, warps_h = 0, take = 0;
    if (range_mode && forced_h == 0) {
      for (int wh = kScoreWarps; wh >= 1; wh /= 2) {
        const int c = 32 * 8 * wh;
        const int padded = (rest + c - 1) / c * c;
        if (padded - rest <= rest / 32 && padded / c <= 2 * num_sms) {
          H = 8; warps_h = wh; take = rest;
          break;
        }
      }
    }
    if (take == 0) {
      if (rest >= 32 * 8 * kScoreWarps) {
        H = 8; warps_h = kScoreWarps; take = rest / (32 * 8 * kScoreWarps) * (32 * 8 * kScoreWarps);
      } else {
        int slots = 32;  // largest power of two <= rest, at least one warp of single hypotheses
        while (slots * 2 <= rest) slots *= 2;
        // round up to the next power of two when that wastes at most a quarter of the launch (or < 64 slots)
        if (rest > slots && 2 * slots - rest <= (slots / 2 > 64 ? slots / 2 : 64)) slots *= 2;
        H = slots / 32 > 8 ? 8 : slots / 32;
        warps_h = slots / (32 * H);
        take = rest < slots ? rest : slots;
      }
    }
    if (forced_h == 4 && H > 4) { warps_h = warps_h * H / 4 > kScoreWarps ? kScoreWarps : warps_h * H / 4; H = 4; }
    const float* X = cloud.x; const float* Y = cloud.y; const float* Z = cloud.z;
    switch (H) {
      case 1: launch_score_h<1>(X, Y, Z, n_per_cloud, cloud_stride, n_clouds, hyps, K, k, k + take, warps_h, t, dot_order, counts, num_sms, s, range_mode, st, chained); break;
      case 2: launch_score_h<2>(X, Y, Z, n_per_cloud, cloud_stride, n_clouds, hyps, K, k, k + take, warps_h, t, dot_order, counts, num_sms, s, range_mode, st, chained); break;
      case 4: launch_score_h<4>(X, Y, Z, n_per_cloud, cloud_stride, n_clouds, hyps, K, k, k + take, warps_h, t, dot_order, counts, num_sms, s, range_mode, st, chained); break;
      default: launch_score_h<8>(X, Y, Z, n_per_cloud, cloud_stride, n_clouds, hyps, K, k, k + take, warps_h, t, dot_order, counts, num_sms, s, range_mode, st, chained); break;
    }
    k += take;
    ++launches;
  }
  return launches;
}


// ------------------------------------------------------------------------------------------------
// Hierarchical scorer (pr_params.scorer = PR_SCORER_HIER): same counts as K2, fewer evaluations.
//
// A copy of the cloud is kept in Morton order; every 32 consecutive points form a block with an
// axis-aligned box (centre c, half-extent e rounded up).  For hypothesis (n, d) and a block,
// |n.p + d| >= |n.c + d| - (|a|ex + |b|ey + |c|ez) for every point of the block, so a block whose box lies
// farther than t + margin from the plane holds no inlier and a block whose box lies entirely within
// t - margin holds only inliers; only the remaining blocks are evaluated point by point, with exactly
// the arithmetic of K2.  margin = 2e-6 * ((|a|+|b|+|c|) * Cmax + |d|) dominates every FP32 rounding error
// of the box test and of the per-point FMA chain (each is at most 3 * 2^-24 of that magnitude), so the
// counts are bit-identical to the brute-force kernel.  Non-finite boxes or hypotheses fail both tests
// and fall through to the exact evaluation.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t morton_spread10(uint32_t v) {
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256) morton_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                     const float* __restrict__ z, size_t n, float lox, float loy, float loz,
                                                     float inv_extent, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float px = x[i], py = y[i], pz = z[i];
    uint32_t k = 0x3FFFFFFFu;  // non-finite points sort to the end
    if (isfinite(px) && isfinite(py) && isfinite(pz)) {
      const uint32_t qx = (uint32_t)fminf(fmaxf((px - lox) * inv_extent, 0.f), 1023.f);
      const uint32_t qy = (uint32_t)fminf(fmaxf((py - loy) * inv_extent, 0.f), 1023.f);
      const uint32_t qz = (uint32_t)fminf(fmaxf((pz - loz) * inv_extent, 0.f), 1023.f);
      k = morton_spread10(qx) | (morton_spread10(qy) << 1) | (morton_spread10(qz) << 2);
    }
    keys[i] = k;
    vals[i] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(256) gather_sorted_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ z, const uint32_t* __restrict__ order,
                                                            size_t n, float* __restrict__ sx, float* __restrict__ sy,
                                                            float* __restrict__ sz, size_t cap) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
    float px = CUDART_NAN_F, py = CUDART_NAN_F, pz = CUDART_NAN_F;
    if (i < n) {
      const uint32_t j = order[i];
      px = x[j];
      py = y[j];
      pz = z[j];
    }
    sx[i] = px;
    sy[i] = py;
    sz[i] = pz;
  }
}

size_t sort_temp_bytes(size_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)n, 0, 30);
  return bytes;
}

// keys/vals: 2 * n uint32 each (in | out halves); temp: sort_temp_bytes(n)
void launch_morton_sort(CloudView src, size_t n, const float lo[3], float extent, uint32_t* keys, uint32_t* vals, void* temp,
                        size_t temp_bytes, CloudView dst, cudaStream_t s) {
  size_t blocks = (dst.cap + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (n) {
    const float inv = extent > 0.f ? 1023.999f / extent : 0.f;
    morton_kernel<<<(unsigned)blocks, 256, 0, s>>>(src.x, src.y, src.z, n, lo[0], lo[1], lo[2], inv, keys, vals);
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys + n, vals, vals + n, (int)n, 0, 30, s);
  }
  gather_sorted_kernel<<<(unsigned)blocks, 256, 0, s>>>(src.x, src.y, src.z, vals + n, n, dst.x, dst.y, dst.z, dst.cap);
}

// one warp per 32-point block: box centre / half-extent (rounded up) over the finite points, their number
__global__ void __launch_bounds__(256) block_bounds_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           const float* __restrict__ z, size_t n_blocks,
                                                           float4* __restrict__ bounds) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_blocks) return;
  const size_t i = warp * 32 + lane;
  const float p[3] = {x[i], y[i], z[i]};  // padding is NaN
  const bool fin = isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]);
  const unsigned fm = __ballot_sync(0xFFFFFFFFu, fin);
  float c[3], e[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const uint32_t k = float_order_key(p[a]);
    const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, fin ? k : 0xFFFFFFFFu);
    const uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, fin ? k : 0u);
    const uint32_t lb = (lo & 0x80000000u) ? (lo ^ 0x80000000u) : ~lo, hb = (hi & 0x80000000u) ? (hi ^ 0x80000000u) : ~hi;
    const float flo = __uint_as_float(lb), fhi = __uint_as_float(hb);
    c[a] = 0.5f * flo + 0.5f * fhi;
    e[a] = fmaxf(__fsub_ru(fhi, c[a]), __fsub_ru(c[a], flo));
  }
  if (lane == 0) {
    if (fm == 0u) {  // no finite point: an infinite box is never culled and never "all in"
      bounds[2 * warp] = make_float4(0.f, 0.f, 0.f, 0.f);
      bounds[2 * warp + 1] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);
    } else {
      bounds[2 * warp] = make_float4(c[0], c[1], c[2], (float)__popc(fm));
      bounds[2 * warp + 1] = make_float4(e[0], e[1], e[2], 0.f);
    }
  }
}

void launch_block_bounds(CloudView sorted, size_t n, float4* bounds, cudaStream_t s) {
  const size_t n_blocks = (n + 31) / 32;
  if (n_blocks == 0) return;
  const size_t ctas = (n_blocks * 32 + 255) / 256;
  block_bounds_kernel<<<(unsigned)ctas, 256, 0, s>>>(sorted.x, sorted.y, sorted.z, n_blocks, bounds);
}

// per hypothesis: (t + margin, t - margin)
__global__ void __launch_bounds__(128) hier_prepare_kernel(const float4* __restrict__ hyps, int K, float t, float cmax,
                                                           float2* __restrict__ aux) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float4 h = hyps[k];
  const float m = __fmaf_ru(__fadd_ru(__fadd_ru(fabsf(h.x), fabsf(h.y)), fabsf(h.z)), cmax, fabsf(h.w));
  const float margin = __fmul_ru(2e-6f, m);
  aux[k] = make_float2(__fadd_ru(t, margin), __fsub_rd(t, margin));
}

constexpr int kHierMaxK = 4096;  // hypotheses per launch chunk (shared-memory counters, 16-bit queue entries)

// Per tile of 1024 sorted points (32 blocks), hypotheses in phases of CH:
//   A1  lane <-> hypothesis: test the tile's box; most hypotheses miss the whole tile.
//   A2  per surviving hypothesis, lane <-> block: test the 32 block boxes; a hit appends the hypothesis
//       to that block's queue (shared memory), an "all in" adds the block's point count.
//   B   per block queue, 32 queued hypotheses at a time, lane <-> hypothesis: the block's 32 points are
//       broadcast from shared memory and evaluated exactly like K2 (FFMA2 / FSET.BF / IADD3); warps
//       take blocks from a shared ticket so uneven queues balance.
// Phase B therefore runs at brute-force efficiency, but only on the (block, hypothesis) pairs whose box
// straddles the threshold slab (a few per cent on the indoor scenes).
// Queue entries that do not fill a group of 32 are carried into the tile's next phase, so lanes idle only
// in the last phase of a tile.
// Dynamic shared memory: points 12 KB | counters 4 * kHierMaxK | queues 32 * (CH + 32) * 2 | heads.
template <int DOT, int CH>
__global__ void __launch_bounds__(256, CH >= 1024 ? 2 : 3)
    score_hier_kernel(const float* __restrict__ SX, const float* __restrict__ SY, const float* __restrict__ SZ,
                      const float4* __restrict__ bounds, size_t n_blocks, int n_tiles, int tiles_per_cta,
                      const float4* __restrict__ hyps, const float2* __restrict__ aux, int K, float t,
                      int32_t* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float* s_x = reinterpret_cast<float*>(s_raw);
  float* s_y = s_x + kTilePoints;
  float* s_z = s_y + kTilePoints;
  int* s_cnt = reinterpret_cast<int*>(s_z + kTilePoints);
  constexpr int CAP = CH + 32;  // per-block queue capacity: a phase's hits plus the carried remainder
  unsigned short* s_q = reinterpret_cast<unsigned short*>(s_cnt + kHierMaxK);  // [32][CAP]
  int* s_qn = reinterpret_cast<int*>(s_q + 32 * CAP);                          // [32] heads + [1] ticket
  int* s_ticket = s_qn + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_begin = blockIdx.y * kHierMaxK;
  const int k_cnt = min(K - k_begin, kHierMaxK);
  const int tile_begin = blockIdx.x * tiles_per_cta;
  const int tile_end = min(n_tiles, tile_begin + tiles_per_cta);
  if (tile_begin >= tile_end) return;
  for (int i = threadIdx.x; i < k_cnt; i += blockDim.x) s_cnt[i] = 0;
  if (threadIdx.x < 33) s_qn[threadIdx.x] = 0;

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    __syncthreads();  // previous tile fully consumed; counters / queue heads initialised
    {
      const size_t base = (size_t)tile * kTilePoints;  // 256 threads x one float4 per plane
      reinterpret_cast<float4*>(s_x)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(SX + base) + threadIdx.x);
      reinterpret_cast<float4*>(s_y)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(SY + base) + threadIdx.x);
      reinterpret_cast<float4*>(s_z)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(SZ + base) + threadIdx.x);
    }
    // lane <-> block of this tile
    const size_t blk = (size_t)tile * 32 + lane;
    const bool have = blk < n_blocks;
    float4 bc = make_float4(0.f, 0.f, 0.f, 0.f), be = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);
    if (have) {
      bc = __ldg(&bounds[2 * blk]);
      be = __ldg(&bounds[2 * blk + 1]);
    }
    // the tile's box from its block boxes (upper bounds rounded up, lower bounds rounded down)
    float tc[3], te[3];
    {
      const float cs[3] = {bc.x, bc.y, bc.z}, es[3] = {be.x, be.y, be.z};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float lo = have ? __fsub_rd(cs[a], es[a]) : CUDART_INF_F;
        float hi = have ? __fadd_ru(cs[a], es[a]) : -CUDART_INF_F;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo = fminf(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
          hi = fmaxf(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
        }
        tc[a] = 0.5f * lo + 0.5f * hi;
        te[a] = fmaxf(__fsub_ru(hi, tc[a]), __fsub_ru(tc[a], lo));
      }
    }
    __syncthreads();

    for (int cb = 0; cb < k_cnt; cb += CH) {
      const int c_end = min(k_cnt, cb + CH);
      // ---- phase A: warp w takes hypotheses cb + 32 * (w + 8 i) ----
      for (int kb = cb + warp * 32; kb < c_end; kb += 256) {
        const int kl = kb + lane;
        float4 h = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        float2 th = make_float2(CUDART_NAN_F, CUDART_NAN_F);
        if (kl < c_end) {
          h = __ldg(&hyps[k_begin + kl]);
          th = __ldg(&aux[k_begin + kl]);
        }
        // A1: tile box.  A miss needs a decisive test; NaN hypotheses fall through (and hit nothing point-wise)
        const float trc = __fmaf_rn(h.x, tc[0], __fmaf_rn(h.y, tc[1], __fmaf_rn(h.z, tc[2], h.w)));
        const float trho = __fmaf_rn(fabsf(h.x), te[0], __fmaf_rn(fabsf(h.y), te[1], __fmul_rn(fabsf(h.z), te[2])));
        unsigned alive = __ballot_sync(0xFFFFFFFFu, kl < c_end && !((fabsf(trc) - trho) > th.x));
        // A2: surviving hypotheses against the 32 block boxes
        while (alive) {
          const int src = __ffs(alive) - 1;
          alive &= alive - 1;
          const float ha = __shfl_sync(0xFFFFFFFFu, h.x, src), hb = __shfl_sync(0xFFFFFFFFu, h.y, src);
          const float hc = __shfl_sync(0xFFFFFFFFu, h.z, src), hd = __shfl_sync(0xFFFFFFFFu, h.w, src);
          const float t_out = __shfl_sync(0xFFFFFFFFu, th.x, src), t_in = __shfl_sync(0xFFFFFFFFu, th.y, src);
          const float rc = __fmaf_rn(ha, bc.x, __fmaf_rn(hb, bc.y, __fmaf_rn(hc, bc.z, hd)));
          const float rho = __fmaf_rn(fabsf(ha), be.x, __fmaf_rn(fabsf(hb), be.y, __fmul_rn(fabsf(hc), be.z)));
          const bool out = (fabsf(rc) - rho) > t_out;
          const bool ain = (fabsf(rc) + rho) < t_in;
          const int q = kb + src;  // launch-chunk-local hypothesis index
          if (have && ain) {
            atomicAdd(&s_cnt[q], (int)bc.w);
          } else if (have && !out) {
            const int pos = atomicAdd(&s_qn[lane], 1);  // < CAP: at most CH new entries per phase + < 32 carried
            s_q[lane * CAP + pos] = (unsigned short)q;
          }
        }
      }
      __syncthreads();

      // ---- phase B: blocks handed out by ticket; 32 queued hypotheses at a time, lane <-> hypothesis ----
      while (true) {
        int b = 0;
        if (lane == 0) b = atomicAdd(s_ticket, 1);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        if (b >= 32) break;
        const int nq = s_qn[b];
        const bool last_phase = cb + CH >= k_cnt;
        const int n_go = last_phase ? nq : (nq & ~31);  // full groups only, except in the tile's last phase
        const float* bx = s_x + b * 32;
        const float* by = s_y + b * 32;
        const float* bz = s_z + b * 32;
        for (int g = 0; g < n_go; g += 32) {
          const bool valid = g + lane < nq;
          const int q = valid ? (int)s_q[b * CAP + g + lane] : 0;
          float4 hq = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
          if (valid) hq = __ldg(&hyps[k_begin + q]);
          unsigned acc = 0u;
#pragma unroll
          for (int s4 = 0; s4 < 8; ++s4) {
            const float4 x4 = *reinterpret_cast<const float4*>(bx + 4 * s4);
            const float4 y4 = *reinterpret_cast<const float4*>(by + 4 * s4);
            const float4 z4 = *reinterpret_cast<const float4*>(bz + 4 * s4);
            const float2 ra = plane_dot2<DOT>(hq.x, hq.y, hq.z, hq.w, make_float2(x4.x, x4.y), make_float2(y4.x, y4.y),
                                              make_float2(z4.x, z4.y));
            const float2 rb = plane_dot2<DOT>(hq.x, hq.y, hq.z, hq.w, make_float2(x4.z, x4.w), make_float2(y4.z, y4.w),
                                              make_float2(z4.z, z4.w));
            float f0, f1, f2, f3;
            asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f0) : "f"(fabsf(ra.x)), "f"(t));
            asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f1) : "f"(fabsf(ra.y)), "f"(t));
            asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f2) : "f"(fabsf(rb.x)), "f"(t));
            asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f3) : "f"(fabsf(rb.y)), "f"(t));
            acc = acc + __float_as_uint(f0) + __float_as_uint(f1);
            acc = acc + __float_as_uint(f2) + __float_as_uint(f3);
          }
          const int c = (int)(((acc >> 23) * 383u) & 511u);  // <= 32 increments: no wrap
          if (valid && c) atomicAdd(&s_cnt[q], c);
        }
        // carry the remainder (< 32 entries) to the front of the queue for the tile's next phase
        const int rem = last_phase ? 0 : nq - n_go;
        unsigned short keep = 0;
        if (lane < rem) keep = s_q[b * CAP + n_go + lane];
        __syncwarp();
        if (lane < rem) s_q[b * CAP + lane] = keep;
        if (lane == 0) s_qn[b] = rem;
      }
      __syncthreads();
      if (threadIdx.x == 0) *s_ticket = 0;
      __syncthreads();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k_cnt; i += blockDim.x) {
    const int v = s_cnt[i];
    if (v) atomicAdd(&counts[k_begin + i], v);
  }
}

template <int DOT, int CH>
static void launch_score_hier_t(dim3 grid, const float* SX, const float* SY, const float* SZ, const float4* bounds,
                                size_t n_blocks, int n_tiles, int tiles_per_cta, const float4* hyps, const float2* aux,
                                int K, float t, int32_t* counts, cudaStream_t s) {
  const size_t smem = 3 * kTilePoints * sizeof(float) + kHierMaxK * sizeof(int) + 32 * (CH + 32) * sizeof(unsigned short) + 33 * sizeof(int);
  // function attributes are per device (a process may hold contexts on several): set on every launch, it is cheap
  if (cudaFuncSetAttribute(score_hier_kernel<DOT, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return;
  score_hier_kernel<DOT, CH><<<grid, 256, smem, s>>>(SX, SY, SZ, bounds, n_blocks, n_tiles, tiles_per_cta, hyps, aux, K, t, counts);
}

int launch_score_hier(CloudView sorted, size_t n, const float4* bounds, const float4* hyps, float2* aux, int K, float t,
                      float cmax, int dot_order, int32_t* counts, int num_sms, cudaStream_t s) {
  if (K <= 0 || n == 0) return 0;
  hier_prepare_kernel<<<(K + 127) / 128, 128, 0, s>>>(hyps, K, t, cmax, aux);
  const int n_tiles = (int)((n + kTilePoints - 1) / kTilePoints);
  const size_t n_blocks = (n + 31) / 32;
  const int chunks = (K + kHierMaxK - 1) / kHierMaxK;
  static const int ch_knob = [] { const char* e = getenv("PR_HIER_CH"); return e ? atoi(e) : 512; }();  // tuning knob (phase size: 256, 512 or 1024 hypotheses)
  const int per_sm = ch_knob >= 1024 ? 2 : 3;
  int gx = per_sm * num_sms / chunks;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  const int tiles_per_cta = (n_tiles + gx - 1) / gx;
  gx = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
  dim3 grid(gx, chunks);
#define PR_HIER(D, C) launch_score_hier_t<D, C>(grid, sorted.x, sorted.y, sorted.z, bounds, n_blocks, n_tiles, tiles_per_cta, hyps, aux, K, t, counts, s)
  if (dot_order == 1) {
    if (ch_knob >= 1024) PR_HIER(1, 1024); else if (ch_knob >= 512) PR_HIER(1, 512); else PR_HIER(1, 256);
  } else {
    if (ch_knob >= 1024) PR_HIER(0, 1024); else if (ch_knob >= 512) PR_HIER(0, 512); else PR_HIER(0, 256);
  }
#undef PR_HIER
  return 2;
}

// ------------------------------------------------------------------------------------------------
// K3: refit moments.  One predicated pass over the cloud; inliers are quantised to a 2^-s grid about
// the pivot and their first and second moments summed as exact integers (second moments split into
// hi * 2^32 + lo so that 64-bit accumulators cannot overflow).  Integer sums commute, so the result
// does not depend on the thread, block or GPU count.
// ------------------------------------------------------------------------------------------------
constexpr int kRefitQueue = 160;  // per-warp inlier queue: up to 31 carried over + 128 new per iteration

// Moments of one quantised inlier (see the section comment): 16 int64 partial sums per thread.
__device__ __forceinline__ void refit_accumulate(long long acc[16], float x, float y, float z, double px, double py,
                                                 double pz, double scale) {
  const int qx = __double2int_rn(((double)x - px) * scale);  // |q| < 2^30 by construction of the scale
  const int qy = __double2int_rn(((double)y - py) * scale);
  const int qz = __double2int_rn(((double)z - pz) * scale);
  acc[0] += 1;
  acc[1] += qx;
  acc[2] += qy;
  acc[3] += qz;
  const long long pr[6] = {(long long)qx * qx, (long long)qx * qy, (long long)qx * qz,
                           (long long)qy * qy, (long long)qy * qz, (long long)qz * qz};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    acc[4 + 2 * k] += pr[k] >> 32;
    acc[5 + 2 * k] += pr[k] & 0xFFFFFFFFll;
  }
}

// Inliers are ~10 % of the points, so evaluating the FP64/int64 moment code under a divergent branch
// would run it at ~10 % lane utilisation and make the pass instruction-bound.  Instead each warp
// pushes its inliers into a shared-memory queue (ballot + popc ranks) and runs the moment code on 32
// queued points at a time; sums commute, so the queue order is irrelevant.
template <int DOT>
__global__ void __launch_bounds__(256, 4) refit_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                       const float* __restrict__ Z, size_t n,
                                                       const float4* __restrict__ hyps, const int4* __restrict__ sample_pts,
                                                       int model_index, float t, double scale, RefitOut* __restrict__ out,
                                                       size_t cloud_stride, int K, const int32_t* __restrict__ model_idx_arr,
                                                       const double* __restrict__ scale_arr, RoundState* st, ChainTail tail) {
  pdl_wait();  // returns at once unless launched as a programmatic dependent (queued rounds, batch path)
  if (st != nullptr) {  // peel loop without the host: size and winning draw of this round live on the device
    if (st->stop || st->best < 0) return;
    chain_stamp(st, kStampRefit);
    n = (size_t)st->n_local;
    model_index = st->best;
  }
  // batch mode (gridDim.y clouds): cloud c uses hypothesis c * K + model_idx_arr[c] and scale_arr[c]
  if (model_idx_arr != nullptr) {
    const size_t c = blockIdx.y;
    const int mi = model_idx_arr[c];
    if (mi < 0) return;  // no model for this cloud
    X += c * cloud_stride;
    Y += c * cloud_stride;
    Z += c * cloud_stride;
    model_index = (int)(c * (size_t)K) + mi;
    scale = scale_arr[c];
    out += c;
  }
  __shared__ float s_q[8][3][kRefitQueue];
  __shared__ long long s_part[8][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  float* qx = s_q[warp][0];
  float* qy = s_q[warp][1];
  float* qz = s_q[warp][2];

  const float4 h = hyps[model_index];
  const int4 pv = sample_pts[3 * model_index];
  const double px = (double)__int_as_float(pv.x), py = (double)__int_as_float(pv.y), pz = (double)__int_as_float(pv.z);
  long long acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0;
  int queued = 0;  // warp-uniform

  const size_t nvec = (n + 3) / 4;  // the tail of the last vector is NaN padding: never an inlier
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // all lanes of a warp run the same number of iterations (ballots below need the full warp)
  const size_t warp_first = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (size_t vb = warp_first; vb < nvec; vb += stride) {
    const size_t v = vb + lane;
    float xs[4], ys[4], zs[4];
    if (v < nvec) {
      const float4 x4 = __ldg(reinterpret_cast<const float4*>(X) + v);
      const float4 y4 = __ldg(reinterpret_cast<const float4*>(Y) + v);
      const float4 z4 = __ldg(reinterpret_cast<const float4*>(Z) + v);
      xs[0] = x4.x; xs[1] = x4.y; xs[2] = x4.z; xs[3] = x4.w;
      ys[0] = y4.x; ys[1] = y4.y; ys[2] = y4.z; ys[3] = y4.w;
      zs[0] = z4.x; zs[1] = z4.y; zs[2] = z4.z; zs[3] = z4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) xs[e] = ys[e] = zs[e] = CUDART_NAN_F;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float r = plane_dot<DOT>(h.x, h.y, h.z, h.w, xs[e], ys[e], zs[e]);
      const bool in = fabsf(r) < t;
      const unsigned m = __ballot_sync(0xFFFFFFFFu, in);
      if (in) {
        const int pos = queued + __popc(m & lt_mask);
        qx[pos] = xs[e];
        qy[pos] = ys[e];
        qz[pos] = zs[e];
      }
      queued += __popc(m);
    }
    __syncwarp();
    while (queued >= 32) {
      queued -= 32;
      refit_accumulate(acc, qx[queued + lane], qy[queued + lane], qz[queued + lane], px, py, pz, scale);
    }
    __syncwarp();
  }
  if (lane < queued) refit_accumulate(acc, qx[lane], qy[lane], qz[lane], px, py, pz, scale);

#pragma unroll
  for (int i = 0; i < 16; ++i) {
    long long v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) s_part[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    long long v = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s_part[w][threadIdx.x];
    if (v != 0) atomicAdd(reinterpret_cast<unsigned long long*>(&out->m[threadIdx.x]), (unsigned long long)v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out->pivot[0] = __int_as_float(pv.x);
    out->pivot[1] = __int_as_float(pv.y);
    out->pivot[2] = __int_as_float(pv.z);
    out->pivot[3] = 0.0f;
  }
  if (tail.rec != nullptr) {
    // one GPU, host-free loop: the last block to finish turns the summed moments into the round's plane
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(tail.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
      __threadfence();
      long long m[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) m[i] = __ldcg(&out->m[i]);
      const float pivot[3] = {__int_as_float(pv.x), __int_as_float(pv.y), __int_as_float(pv.z)};
      chain_finish(st, hyps, tail.triples, m, pivot, 1, tail.scale_exp, tail.n_draws, tail.rec);
    }
  }
}

void launch_refit(CloudView cloud, size_t n, const float4* hyps, const int4* sample_pts, int model_index, float t,
                  int dot_order, int scale_exp, RefitOut* out, int num_sms, cudaStream_t s, RoundState* st, const ChainTail* tail) {
  const double scale = ldexp(1.0, scale_exp);
  ChainTail tl;
  if (tail) tl = *tail;
  size_t nvec = (n + 3) / 4;
  size_t blocks = (nvec + 255) / 256;
  // one wave exactly: 4 resident CTAs per SM (launch bounds), grid-stride over the cloud
  if (blocks > (size_t)num_sms * 4) blocks = (size_t)num_sms * 4;
  if (blocks < 1) blocks = 1;
  if (st != nullptr) {
    if (dot_order == 1)
      launch_chained(refit_kernel<1>, dim3((unsigned)blocks), dim3(256), 0, s, cloud.x, cloud.y, cloud.z, n, hyps, sample_pts, model_index, t, scale, out,
                     (size_t)0, 0, (const int32_t*)nullptr, (const double*)nullptr, st, tl);
    else
      launch_chained(refit_kernel<0>, dim3((unsigned)blocks), dim3(256), 0, s, cloud.x, cloud.y, cloud.z, n, hyps, sample_pts, model_index, t, scale, out,
                     (size_t)0, 0, (const int32_t*)nullptr, (const double*)nullptr, st, tl);
  } else if (dot_order == 1)
    refit_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(cloud.x, cloud.y, cloud.z, n, hyps, sample_pts, model_index, t, scale, out, 0, 0, nullptr, nullptr, st, tl);
  else
    refit_kernel<0><<<(unsigned)blocks, 256, 0, s>>>(cloud.x, cloud.y, cloud.z, n, hyps, sample_pts, model_index, t, scale, out, 0, 0, nullptr, nullptr, st, tl);
}

void launch_refit_batch(CloudView clouds, size_t n_per, size_t cloud_stride, int n_clouds, const float4* hyps,
                        const int4* sample_pts, int K, const int32_t* model_idx, float t, int dot_order,
                        const double* scales, RefitOut* outs, cudaStream_t s, bool chained) {
  if (n_clouds <= 0) return;
  size_t nvec = (n_per + 3) / 4;
  // blocks per cloud: enough to fill the machine about twice, no more — every block ends with a 16 x int64 reduction
  // (160 shuffles per warp), which costs as much as ~4 loop iterations (512 clouds of 32K points, B200: 86.5 us with 8
  // blocks per cloud, 74.6 with 2)
  unsigned bx = (unsigned)((nvec + 255) / 256);
  unsigned bx_max = (unsigned)((2 * 4 * 148 + n_clouds - 1) / n_clouds);
  if (bx_max > 8) bx_max = 8;
  if (bx > bx_max) bx = bx_max;
  if (bx < 1) bx = 1;
  dim3 grid(bx, (unsigned)n_clouds);
  if (chained) {
    if (dot_order == 1)
      launch_chained(refit_kernel<1>, grid, dim3(256), 0, s, clouds.x, clouds.y, clouds.z, n_per, hyps, sample_pts, 0, t, 1.0, outs, cloud_stride, K,
                     model_idx, scales, (RoundState*)nullptr, ChainTail());
    else
      launch_chained(refit_kernel<0>, grid, dim3(256), 0, s, clouds.x, clouds.y, clouds.z, n_per, hyps, sample_pts, 0, t, 1.0, outs, cloud_stride, K,
                     model_idx, scales, (RoundState*)nullptr, ChainTail());
  } else if (dot_order == 1)
    refit_kernel<1><<<grid, 256, 0, s>>>(clouds.x, clouds.y, clouds.z, n_per, hyps, sample_pts, 0, t, 1.0, outs, cloud_stride, K, model_idx, scales, nullptr, ChainTail());
  else
    refit_kernel<0><<<grid, 256, 0, s>>>(clouds.x, clouds.y, clouds.z, n_per, hyps, sample_pts, 0, t, 1.0, outs, cloud_stride, K, model_idx, scales, nullptr, ChainTail());
}

// ------------------------------------------------------------------------------------------------
// PCL 1.8's own refit accumulation (PR_REFIT_PCL_FLOAT): computeMeanAndCovarianceMatrix sums nine FP32 accumulators
// over the inliers sequentially, in index order.  A sequential FP32 sum cannot be split across threads without
// changing its value, so one thread (lane 0) does all the additions; the other lanes of its warp only fetch the next
// 32 inliers' coordinates (a coalesced index read + gather) and hand them over by shuffle.  A parity mode, not the fast
// path: ~20 ms per million inliers.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) refit_pcl_float_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Z,
                                                             const int32_t* __restrict__ idx, const long long* __restrict__ n_idx_dev,
                                                             float* __restrict__ out9) {
  const long long n_idx = *n_idx_dev;
  const int lane = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f, a8 = 0.f;
  for (long long base = 0; base < n_idx; base += 32) {
    float x = 0.f, y = 0.f, z = 0.f;
    if (base + lane < n_idx) {
      const int32_t j = idx[base + lane];
      x = X[j]; y = Y[j]; z = Z[j];
    }
    const int m = (int)(n_idx - base < 32 ? n_idx - base : 32);
    for (int k = 0; k < m; ++k) {
      const float px = __shfl_sync(0xFFFFFFFFu, x, k), py = __shfl_sync(0xFFFFFFFFu, y, k), pz = __shfl_sync(0xFFFFFFFFu, z, k);
      a0 = __fadd_rn(a0, __fmul_rn(px, px));
      a1 = __fadd_rn(a1, __fmul_rn(px, py));
      a2 = __fadd_rn(a2, __fmul_rn(px, pz));
      a3 = __fadd_rn(a3, __fmul_rn(py, py));
      a4 = __fadd_rn(a4, __fmul_rn(py, pz));
      a5 = __fadd_rn(a5, __fmul_rn(pz, pz));
      a6 = __fadd_rn(a6, px);
      a7 = __fadd_rn(a7, py);
      a8 = __fadd_rn(a8, pz);
    }
  }
  if (lane == 0) {
    out9[0] = a0; out9[1] = a1; out9[2] = a2; out9[3] = a3; out9[4] = a4; out9[5] = a5; out9[6] = a6; out9[7] = a7; out9[8] = a8;
  }
}

void launch_refit_pcl_float(CloudView cloud, const int32_t* idx, const long long* n_idx_dev, float* out9, cudaStream_t s) {
  refit_pcl_float_kernel<<<1, 32, 0, s>>>(cloud.x, cloud.y, cloud.z, idx, n_idx_dev, out9);
}

// ------------------------------------------------------------------------------------------------
// K5: stable partition with a single-pass decoupled look-back scan.
// Tile = 2048 points (256 threads x 2 rounds x 4 consecutive points, 128-bit loads).  The scan runs
// on the number of KEPT points; the inlier rank of a point is its position minus the kept points
// before it, so one scan serves both outputs.  Both outputs are staged in shared memory (kept points
// from the front, inliers from the back of the same arrays) and written out contiguously.
//
// The look-back is block-wide: thread j inspects the descriptor of tile - 1 - j, so one step covers 256
// predecessors.  With ~740 tiles in flight on 148 SMs the nearest tile whose inclusive prefix is already
// published lies hundreds of tiles back; a warp-wide (32-descriptor) window walks that distance in ~20
// dependent L2 round trips per tile, which is what bounded the round-1 kernel at 66 % of the HBM rate
// (tile lifetime ~10 us against ~2 us of loads and stores); 256 descriptors per step make it one or two.
// ------------------------------------------------------------------------------------------------
constexpr int kCompactThreads = 256;
constexpr unsigned long long kTileAggregate = 1ull << 62;
constexpr unsigned long long kTileInclusive = 2ull << 62;
constexpr unsigned long long kTileValueMask = (1ull << 62) - 1;

template <int DOT, bool WRITE_REM>
__global__ void __launch_bounds__(kCompactThreads, 5)
    compact_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Z,
                   const int32_t* __restrict__ O, size_t n, Plane4 pl, float t, float* __restrict__ DX,
                   float* __restrict__ DY, float* __restrict__ DZ, int32_t* __restrict__ DO, size_t dst_cap,
                   int32_t* __restrict__ inl_cur, int32_t* __restrict__ inl_orig, unsigned long long* tile_state,
                   unsigned* ticket, long long* __restrict__ totals, const uint32_t* __restrict__ flags,
                   RoundState* st, ChainTail tail) {
  __shared__ __align__(16) float s_x[kCompactTile];
  __shared__ __align__(16) float s_y[kCompactTile];
  __shared__ __align__(16) float s_z[kCompactTile];
  __shared__ __align__(16) int s_o[kCompactTile];
  __shared__ int s_warp[2][8];
  __shared__ unsigned s_tile;
  __shared__ long long s_lb_sum[8];
  __shared__ int s_lb_found[8];

  if (st != nullptr) {
    // peel loop without the host: size, plane and list offset of this round live on the device; the grid was sized
    // for an upper bound of n
    pdl_wait();
    if (st->stop) return;
    chain_stamp(st, kStampPeel);
    n = (size_t)st->n_local;
    pl.a = st->plane[0]; pl.b = st->plane[1]; pl.c = st->plane[2]; pl.d = st->plane[3];
    if (inl_cur) inl_cur += st->inl_off;
    if (inl_orig) inl_orig += st->inl_off;
  }
  const unsigned n_tiles = (unsigned)((n + kCompactTile - 1) / kCompactTile);

  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  if (n_tiles == 0) {  // empty cloud (a shard whose points were all peeled): totals and padding only
    if (tile == 0) {
      if (threadIdx.x == 0) {
        totals[0] = 0;
        totals[1] = 0;
        if (tail.rec != nullptr) chain_advance(st, totals, 1, 0, tail.min_plane, tail.rec);
      }
      if (WRITE_REM) {
        size_t pad_end = kTilePoints;
        if (pad_end > dst_cap) pad_end = dst_cap;
        for (size_t j = threadIdx.x; j < pad_end; j += kCompactThreads) { DX[j] = CUDART_NAN_F; DY[j] = CUDART_NAN_F; DZ[j] = CUDART_NAN_F; }
      }
    }
    return;
  }
  if (tile >= n_tiles) return;  // surplus CTAs of an upper-bound grid
  const size_t base = (size_t)tile * kCompactTile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  float xs[2][4], ys[2][4], zs[2][4];
  int os[2][4];
  unsigned keep[2], inl[2];
  int kc[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const size_t i0 = base + (size_t)u * 1024 + 4 * threadIdx.x;
    const float4 x4 = __ldg(reinterpret_cast<const float4*>(X + i0));
    const float4 y4 = __ldg(reinterpret_cast<const float4*>(Y + i0));
    const float4 z4 = __ldg(reinterpret_cast<const float4*>(Z + i0));
    xs[u][0] = x4.x; xs[u][1] = x4.y; xs[u][2] = x4.z; xs[u][3] = x4.w;
    ys[u][0] = y4.x; ys[u][1] = y4.y; ys[u][2] = y4.z; ys[u][3] = y4.w;
    zs[u][0] = z4.x; zs[u][1] = z4.y; zs[u][2] = z4.z; zs[u][3] = z4.w;
    if (O != nullptr) {
      const int4 o4 = __ldg(reinterpret_cast<const int4*>(O + i0));
      os[u][0] = o4.x; os[u][1] = o4.y; os[u][2] = o4.z; os[u][3] = o4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) os[u][e] = (int)(i0 + e);
    }
    keep[u] = 0u;
    inl[u] = 0u;
    uint4 f4 = make_uint4(0u, 0u, 0u, 0u);
    if (DOT == 3) f4 = __ldg(reinterpret_cast<const uint4*>(flags + i0));
    const unsigned fl[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool valid = (i0 + e) < n;
      bool in;
      if (DOT == 2) {  // staging filter (pcl::removeNaNFromPointCloud): "inliers" are the non-finite points
        in = !(isfinite(xs[u][e]) && isfinite(ys[u][e]) && isfinite(zs[u][e]));
      } else if (DOT == 3) {  // postProcessPlanes: "inliers" are the points some plane polygon claimed
        in = fl[e] != 0u;
      } else {
        const float r = plane_dot<DOT>(pl.a, pl.b, pl.c, pl.d, xs[u][e], ys[u][e], zs[u][e]);
        in = fabsf(r) < t;
      }
      if (valid && !in) keep[u] |= 1u << e;
      if (valid && in) inl[u] |= 1u << e;
    }
    kc[u] = __popc(keep[u]);
  }

  // block-wide exclusive scan of the kept counts, per round
  int lane_excl[2], round_total[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int v = kc[u];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xFFFFFFFFu, v, o);
      if (lane >= o) v += y;
    }
    lane_excl[u] = v - kc[u];
    if (lane == 31) s_warp[u][warp] = v;
  }
  __syncthreads();
  int warp_excl[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int run = 0, mine = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w == warp) mine = run;
      run += s_warp[u][w];
    }
    warp_excl[u] = mine;
    round_total[u] = run;
  }
  const int keep_total = round_total[0] + round_total[1];
  const size_t remaining_pts = n - base;
  const int valid_in_tile = remaining_pts < (size_t)kCompactTile ? (int)remaining_pts : kCompactTile;
  const int inl_total = valid_in_tile - keep_total;

  // publish this tile's aggregate before anything waits on a predecessor
  if (threadIdx.x == 0)
    atomicExch(&tile_state[tile], (tile == 0 ? kTileInclusive : kTileAggregate) | (unsigned long long)keep_total);

  // stage both partitions in shared memory (local ranks only: overlaps the predecessors' progress)
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int kr = (u ? round_total[0] : 0) + warp_excl[u] + lane_excl[u];  // kept rank within the tile
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int lp = u * 1024 + 4 * threadIdx.x + e;  // local position
      if (keep[u] & (1u << e)) {
        if (WRITE_REM) {
          s_x[kr] = xs[u][e];
          s_y[kr] = ys[u][e];
          s_z[kr] = zs[u][e];
          s_o[kr] = os[u][e];
        }
        ++kr;
      } else if (inl[u] & (1u << e)) {
        const int ir = lp - kr;  // inlier rank within the tile
        s_x[kCompactTile - 1 - ir] = __int_as_float(lp);
        s_o[kCompactTile - 1 - ir] = os[u][e];
      }
    }
  }

  // block-wide decoupled look-back: kept points of all earlier tiles
  long long excl = 0;
  if (tile == 0) {
    __syncthreads();
  } else {
    long long back = (long long)tile - 1;
    while (true) {
      const long long idx = back - (long long)threadIdx.x;
      unsigned long long stv = kTileInclusive;  // positions before tile 0 contribute an inclusive 0
      if (idx >= 0) {
        unsigned spins = 0;
        do {
          stv = *reinterpret_cast<volatile unsigned long long*>(&tile_state[idx]);
          if (++spins > (1u << 24)) __trap();  // seconds of polling: a lost descriptor becomes a launch failure, not a hung GPU
        } while ((stv >> 62) == 0ull);
      }
      const unsigned incl_mask = __ballot_sync(0xFFFFFFFFu, (stv >> 62) == 2ull);
      const int first_incl = incl_mask ? (__ffs(incl_mask) - 1) : 32;
      long long val = (lane <= first_incl) ? (long long)(stv & kTileValueMask) : 0ll;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xFFFFFFFFu, val, o);
      if (lane == 0) {
        s_lb_sum[warp] = val;
        s_lb_found[warp] = incl_mask != 0u;
      }
      __syncthreads();
      bool found = false;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        if (!found) {
          excl += s_lb_sum[w];
          found = s_lb_found[w] != 0;
        }
      }
      if (found) break;
      back -= kCompactThreads;
      __syncthreads();  // the per-warp slots are rewritten by the next step
    }
    if (threadIdx.x == 0) atomicExch(&tile_state[tile], kTileInclusive | (unsigned long long)(excl + keep_total));
  }
  const long long inl_excl = (long long)base - excl;

  if (WRITE_REM) {
    for (int j = threadIdx.x; j < keep_total; j += kCompactThreads) {
      DX[excl + j] = s_x[j];
      DY[excl + j] = s_y[j];
      DZ[excl + j] = s_z[j];
      if (DO != nullptr) DO[excl + j] = s_o[j];
    }
  }
  for (int j = threadIdx.x; j < inl_total; j += kCompactThreads) {
    const int lp = __float_as_int(s_x[kCompactTile - 1 - j]);
    if (inl_cur) inl_cur[inl_excl + j] = (int32_t)(base + lp);
    if (inl_orig) inl_orig[inl_excl + j] = s_o[kCompactTile - 1 - j];
  }

  if (tile == n_tiles - 1) {
    const long long total_keep = excl + keep_total;
    if (threadIdx.x == 0) {
      totals[0] = total_keep;
      totals[1] = (long long)n - total_keep;
      // one GPU, host-free loop: the stop rule and the next round's sizes.  Every CTA that owns a tile read the state
      // before it took its ticket, i.e. before this (last) tile could complete its look-back; surplus CTAs that start
      // later see a smaller cloud or the stop flag and leave either way.
      if (tail.rec != nullptr) chain_advance(st, totals, 1, 0, tail.min_plane, tail.rec);
    }
    if (WRITE_REM) {
      size_t pad_end = ((size_t)total_keep + kTilePoints - 1) / kTilePoints * kTilePoints + kTilePoints;
      if (pad_end > dst_cap) pad_end = dst_cap;
      for (size_t j = (size_t)total_keep + threadIdx.x; j < pad_end; j += kCompactThreads) {
        DX[j] = CUDART_NAN_F;
        DY[j] = CUDART_NAN_F;
        DZ[j] = CUDART_NAN_F;
      }
    }
  }
}

size_t compact_scratch_bytes(size_t n) {
  size_t n_tiles = (n + kCompactTile - 1) / kCompactTile;
  return (n_tiles + 2) * sizeof(unsigned long long);
}

void launch_compact(CloudView src, size_t n, Plane4 plane, float t, int dot_order, CloudView dst, bool write_remaining,
                    int32_t* inl_cur, int32_t* inl_orig, void* scratch, long long* totals, cudaStream_t s, const uint32_t* flags,
                    RoundState* st, const ChainTail* tail) {
  if (n == 0 && st == nullptr) {
    cudaMemsetAsync(totals, 0, 2 * sizeof(long long), s);
    return;
  }
  unsigned n_tiles = (unsigned)((n + kCompactTile - 1) / kCompactTile);
  if (n_tiles == 0) n_tiles = 1;  // st mode: one CTA writes the totals and the padding of an empty shard
  if (st == nullptr) cudaMemsetAsync(scratch, 0, compact_scratch_bytes(n), s);  // st mode: cleared by the round's prep kernel
  ChainTail tl;
  if (tail) tl = *tail;
  unsigned long long* tile_state = reinterpret_cast<unsigned long long*>(scratch);
  unsigned* ticket = reinterpret_cast<unsigned*>(tile_state + (n + kCompactTile - 1) / kCompactTile);
#define PR_COMPACT(D, W)                                                                                                       \
  do {                                                                                                                         \
    if (st != nullptr)                                                                                                         \
      launch_chained(compact_kernel<D, W>, dim3(n_tiles), dim3(kCompactThreads), 0, s, src.x, src.y, src.z, src.orig, n, plane, t, dst.x, \
                     dst.y, dst.z, dst.orig, dst.cap, inl_cur, inl_orig, tile_state, ticket, totals, flags, st, tl);           \
    else                                                                                                                       \
      compact_kernel<D, W><<<n_tiles, kCompactThreads, 0, s>>>(src.x, src.y, src.z, src.orig, n, plane, t, dst.x, dst.y, dst.z, \
                                                               dst.orig, dst.cap, inl_cur, inl_orig, tile_state, ticket,       \
                                                               totals, flags, st, tl);                                         \
  } while (0)
  if (dot_order == 3) {
    PR_COMPACT(3, true);
  } else if (dot_order == 2) {
    PR_COMPACT(2, true);
  } else if (dot_order == 1) {
    if (write_remaining) PR_COMPACT(1, true); else PR_COMPACT(1, false);
  } else {
    if (write_remaining) PR_COMPACT(0, true); else PR_COMPACT(0, false);
  }
#undef PR_COMPACT
}

// ------------------------------------------------------------------------------------------------
// measurement helpers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
  float2 r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = make_float2(threadIdx.x * 0.001f + i, (float)i);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = __ffma2_rn(r[i], A, B);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i].x + r[i].y;
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

void launch_ffma_peak(float* out, int iters, int grid, cudaStream_t s) {
  ffma_peak_kernel<<<grid, 256, 0, s>>>(out, iters, 1.0001f, 0.5f);
}

__global__ void __launch_bounds__(256) copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __ldg(&src[i]);
}

__global__ void __launch_bounds__(256) fill_kernel(float4* __restrict__ dst, size_t n, float v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = make_float4(v, v, v, v);
}

void launch_copy(const float4* src, float4* dst, size_t n_vec, int num_sms, cudaStream_t s) {
  copy_kernel<<<num_sms * 8, 256, 0, s>>>(src, dst, n_vec);
}

void launch_fill(float4* dst, size_t n_vec, float v, int num_sms, cudaStream_t s) {
  fill_kernel<<<num_sms * 8, 256, 0, s>>>(dst, n_vec, v);
}

}  // namespace pr
