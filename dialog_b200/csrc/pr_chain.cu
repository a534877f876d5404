// pr_chain.cu — the small kernels that take the host out of the peel loop (score-all mode).
//
// One round of pcl::SACSegmentation::segment + ExtractIndices (the reference's peel at Dialog/PlaneDetect.h:1560-1566)
// is a fixed kernel sequence whose sizes and decisions live in a RoundState in HBM:
//   draw_scatter / draw_resolve   PCL's index triples for the round's cloud size (pr_draw.h)
//   gather + models + score       K1, K2 (pr_kernels.cu), sized from RoundState on the device
//   replay_kernel                 RandomSampleConsensus::computeModel's decision over the K counts
//   refit (+ chain_finish)        K3 moments; its last block solves pcl::eigen33's closed form (pr_math.h) -> refined plane
//   compact (+ chain_advance)     K5 peel; its last tile applies the minimum-plane-size rule and sets the next round's sizes
// so the host queues whole rounds ahead and only reads the per-round records (pr_api.cpp run_chain).
#include "pr_kernels.h"

#include <math_constants.h>

#include "pr_chain_dev.cuh"
#include "pr_draw.h"
#include "pr_math.h"

namespace pr {

namespace {
struct DevAtomics {
  static __device__ __forceinline__ unsigned long long cas(unsigned long long* p, unsigned long long expect, unsigned long long v) {
    return atomicCAS(p, expect, v);
  }
  static __device__ __forceinline__ unsigned long long load(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
  }
  static __device__ __forceinline__ uint32_t add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
};
}  // namespace

// ---- sampler, parallel phase: one thread per op; the same launch clears everything the round accumulates into (none of
// which it reads itself) ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) draw_scatter_kernel(const uint32_t* __restrict__ rnd, int n_ops, const RoundState* __restrict__ st,
                                                           int32_t* __restrict__ v, unsigned long long* table, uint32_t table_mask,
                                                           uint32_t epoch, uint32_t* coll, uint32_t* coll_count,
                                                           int32_t* __restrict__ counts, int K, RefitOut* __restrict__ refit,
                                                           unsigned long long* __restrict__ scratch, size_t scratch_words,
                                                           unsigned* __restrict__ tickets) {
  pdl_wait();
  if (st->stop) return;
  chain_stamp(st, kStampDraw);
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = i0; i < scratch_words; i += stride) scratch[i] = 0ull;
  for (size_t i = i0; i < (size_t)K; i += stride) counts[i] = 0;
  if (i0 < sizeof(RefitOut) / sizeof(long long)) reinterpret_cast<long long*>(refit)[i0] = 0;
  if (i0 == 0) tickets[0] = 0u;
  if (st->n_global < 3) return;
  const uint32_t n_points = (uint32_t)st->n_global;
  for (int s = (int)i0; s < n_ops; s += (int)stride)
    draw_scatter<DevAtomics>((uint32_t)s, rnd[s], n_points, v, table, table_mask, epoch, coll, coll_count, (uint32_t)kDrawCollCap);
}

// ---- sampler, sequential phase: the colliding ops, sorted by op index, replayed by one thread ------------------------------
// Everything the replay reads is staged in shared memory first (the sorted list, v at each op and at the three ops before
// it, the position map), so the sequential part never waits on HBM.
constexpr int kResolveThreads = 1024;
constexpr int kResolveMapSlots = 2 * kDrawMaxCollisions;
constexpr size_t kResolveSmemBytes = (size_t)kDrawCollCap * 4 + (size_t)kDrawCollCap * 16 + (size_t)kResolveMapSlots * 8;

struct SmemFetch {
  const uint32_t* sorted;
  const int32_t* pre;  // 4 per listed op: v[s], v[s - 1], v[s - 2], v[s - 3]
  __device__ __forceinline__ int32_t operator()(int i, uint32_t idx) const { return pre[4 * i + (int)(sorted[i] - idx)]; }
};

__global__ void __launch_bounds__(kResolveThreads) draw_resolve_kernel(RoundState* st, int32_t* v, const uint32_t* __restrict__ coll,
                                                                       uint32_t* coll_count, RoundRecord* rec) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint32_t* s_sorted = reinterpret_cast<uint32_t*>(s_raw);
  int32_t* s_pre = reinterpret_cast<int32_t*>(s_raw + (size_t)kDrawCollCap * 4);
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(s_raw + (size_t)kDrawCollCap * 20);
  int32_t* s_vals = reinterpret_cast<int32_t*>(s_raw + (size_t)kDrawCollCap * 20 + (size_t)kResolveMapSlots * 4);
  __shared__ int s_distinct;
  pdl_wait();
  if (st->stop) return;
  chain_stamp(st, kStampResolve);
  if (st->n_global < 3) {  // getSamples: "Can not select 0 unique points out of N": segment() returns no model
    if (threadIdx.x == 0) {
      st->stop = 1;
      st->best = -1;
      rec->stop = 1;
      rec->n_cloud = st->n_global;
      rec->n_local = st->n_local;
      rec->n_rem_local = st->n_local;
      rec->n_rem_global = st->n_global;
      rec->first_after = st->first;
      rec->inl_off = st->inl_off;
      __threadfence_system();
      rec->ran = 1;
    }
    return;
  }
  const uint32_t n_c = *coll_count;
  __syncthreads();
  if (threadIdx.x == 0) *coll_count = 0u;  // ready for the next round's scatter (the host zeroes it before the first)
  if (n_c > (uint32_t)kDrawCollCap) {  // crowded round: the sequential host sampler takes it
    if (threadIdx.x == 0) {
      st->stop = 2;
      rec->stop = 2;
      __threadfence_system();
      rec->ran = 1;
    }
    return;
  }
  for (int i = threadIdx.x; i < kResolveMapSlots; i += kResolveThreads) s_keys[i] = kDrawNoOp;
  if (threadIdx.x == 0) s_distinct = 0;
  // rank sort (the list is tens of entries for the clouds this path is used on)
  for (uint32_t i = threadIdx.x; i < n_c; i += kResolveThreads) {
    const uint32_t mine = coll[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n_c; ++j) {
      const uint32_t o = coll[j];
      rank += (o < mine || (o == mine && j < i)) ? 1u : 0u;
    }
    s_sorted[rank] = mine;
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n_c; i += kResolveThreads) {
    const uint32_t s = s_sorted[i];
    if (i == 0 || s != s_sorted[i - 1]) atomicAdd(&s_distinct, 1);
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) s_pre[4 * i + k] = s >= k ? v[s - k] : 0;
  }
  __syncthreads();
  SmemFetch fetch{s_sorted, s_pre};
  // common case: no listed op swaps inside the head -> every entry follows its own dependency chain, one thread each (pr_draw.h)
  int ok = 1;
  for (uint32_t i = threadIdx.x; i < n_c; i += kResolveThreads) ok = ok && draw_independent_ok(s_sorted, (int)i, fetch);
  if (__syncthreads_and(ok)) {
    for (uint32_t i = threadIdx.x; i < n_c; i += kResolveThreads) v[s_sorted[i]] = draw_resolve_independent(s_sorted, (int)i, fetch);
    return;
  }
  if (threadIdx.x != 0) return;
  if (s_distinct > kDrawMaxCollisions) {
    st->stop = 2;
    rec->stop = 2;
    __threadfence_system();
    rec->ran = 1;
    return;
  }
  draw_resolve(s_sorted, (int)n_c, v, fetch, s_keys, s_vals, (uint32_t)(kResolveMapSlots - 1));
}

// ---- computeModel's decision over the K counts (pr_chain_dev.cuh chain_replay_block) as its own kernel (one GPU) ------
__global__ void __launch_bounds__(kChainBlock) replay_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ good, int K,
                                                             RoundState* st, RoundRecord* rec) {
  pdl_wait();
  if (st->stop) return;
  chain_stamp(st, kStampDecide);
  chain_replay_block(counts, good, K, st, rec);
}

// ---- refined plane as its own kernel (rounds without a refit pass: optimize_coefficients off; with the refit it runs in
// the last block of refit_kernel on one GPU, inside the moments exchange when sharded; the stop rule likewise runs in
// the last tile of compact_kernel / inside the totals exchange) ---------------------------------------------------------
__global__ void finish_kernel(RoundState* st, const float4* __restrict__ hyps, const int32_t* __restrict__ triples,
                              const RefitOut* __restrict__ refit, int optimize, int scale_exp, int n_draws, RoundRecord* rec) {
  pdl_wait();
  if (threadIdx.x != 0 || st->stop) return;
  chain_stamp(st, kStampFinish);
  long long m[16];
  for (int i = 0; i < 16; ++i) m[i] = refit->m[i];
  const float pivot[3] = {refit->pivot[0], refit->pivot[1], refit->pivot[2]};
  chain_finish(st, hyps, triples, m, pivot, optimize, scale_exp, n_draws, rec);
}

void launch_draw(const uint32_t* rnd, int n_draws, RoundState* st, int32_t* triples, unsigned long long* table, size_t table_slots,
                 uint32_t epoch, uint32_t* coll, uint32_t* coll_count, RoundRecord* rec, int32_t* counts, RefitOut* refit, void* scratch,
                 size_t scratch_bytes, unsigned* tickets, cudaStream_t s) {
  const int n_ops = 3 * n_draws;
  int blocks = (n_ops + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  launch_chained(draw_scatter_kernel, dim3(blocks), dim3(256), 0, s, rnd, n_ops, st, triples, table, (uint32_t)(table_slots - 1), epoch, coll,
                 coll_count, counts, n_draws, refit, reinterpret_cast<unsigned long long*>(scratch), scratch_bytes / 8, tickets);
  cudaFuncSetAttribute(draw_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResolveSmemBytes);  // per device: cheap, unconditional
  launch_chained(draw_resolve_kernel, dim3(1), dim3(kResolveThreads), kResolveSmemBytes, s, st, triples, coll, coll_count, rec);
}

size_t draw_table_slots(int n_draws) {
  size_t cap = 1024;
  while (cap < 12 * (size_t)n_draws) cap <<= 1;
  return cap;
}

void launch_replay(const int32_t* counts, const int32_t* good, int K, RoundState* st, RoundRecord* rec, cudaStream_t s) {
  launch_chained(replay_kernel, dim3(1), dim3(kChainBlock), 0, s, counts, good, K, st, rec);
}

void launch_finish(RoundState* st, const float4* hyps, const int32_t* triples, const RefitOut* refit, int optimize, int scale_exp,
                   int n_draws, RoundRecord* rec, cudaStream_t s) {
  launch_chained(finish_kernel, dim3(1), dim3(32), 0, s, st, hyps, triples, refit, optimize, scale_exp, n_draws, rec);
}

}  // namespace pr

// =====================================================================================================================
// Batch of equal-sized small clouds (BASELINE configs[4]: per-scan tiles), score-all mode, without the host in the loop:
// one segment() per cloud — computeModel's decision, the closed-form refit and the final selection with its index list
// (pcl::SACSegmentation::segment's `inliers`) — as a fixed sequence of launches over all clouds.
// =====================================================================================================================
namespace pr {

// computeModel over the K counts of every cloud: best[c] = first draw with the largest count, or -1 and *any_bad = 1 when
// a degenerate sample among the K draws means PCL would have drawn further (the host-driven path then redoes the batch).
__global__ void __launch_bounds__(128) batch_replay_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ good, int K,
                                                           int32_t* __restrict__ best, int32_t* __restrict__ best_count, int* any_bad,
                                                           RefitOut* __restrict__ refit_to_clear) {
  __shared__ unsigned long long s_best[4];
  pdl_wait();
  const size_t c = blockIdx.x;
  if (refit_to_clear != nullptr && threadIdx.x < sizeof(RefitOut) / sizeof(long long))
    reinterpret_cast<long long*>(refit_to_clear + c)[threadIdx.x] = 0;  // the refit pass that follows accumulates into it
  unsigned long long b = 0ull;
  int all_good = 1;
  for (int j = threadIdx.x; j < K; j += 128) {
    if (!good[c * K + j]) all_good = 0;
    const unsigned long long key = ((unsigned long long)(unsigned)counts[c * K + j] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)j);
    b = key > b ? key : b;
  }
  all_good = __syncthreads_and(all_good);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_down_sync(0xFFFFFFFFu, b, o);
    b = other > b ? other : b;
  }
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int w = 1; w < 4; ++w) b = s_best[w] > b ? s_best[w] : b;
  if (!all_good) {
    best[c] = -1;
    best_count[c] = 0;
    *any_bad = 1;
    return;
  }
  best[c] = (int32_t)(0xFFFFFFFFu - (unsigned)(b & 0xFFFFFFFFull));
  best_count[c] = (int32_t)(b >> 32);
}

// refined[c] = closed-form plane from cloud c's moments (or its raw model), thread per cloud
__global__ void __launch_bounds__(64) batch_finish_kernel(const float4* __restrict__ hyps, int K, const int32_t* __restrict__ best,
                                                          const RefitOut* __restrict__ refit, const int32_t* __restrict__ scale_exp,
                                                          int optimize, int n_clouds, float4* __restrict__ raw, float4* __restrict__ refined) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clouds) return;
  const int b = best[c];
  if (b < 0) {
    raw[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    refined[c] = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);  // never an inlier: the final count is 0
    return;
  }
  const float4 h = hyps[(size_t)c * K + b];
  float out[4] = {h.x, h.y, h.z, h.w};
  if (optimize) {
    long long m[16];
    for (int i = 0; i < 16; ++i) m[i] = refit[c].m[i];
    const float pivot[3] = {refit[c].pivot[0], refit[c].pivot[1], refit[c].pivot[2]};
    pm_plane_from_moments(m, pivot, scale_exp[c], out);
  }
  raw[c] = h;
  refined[c] = make_float4(out[0], out[1], out[2], out[3]);
}

// offs[0 .. C] = exclusive scan of cnt[0 .. C) (one block; C is a few thousand)
__global__ void __launch_bounds__(1024) batch_offsets_kernel(const int32_t* __restrict__ cnt, int n_clouds, unsigned long long* __restrict__ offs) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0ull;
  pdl_wait();
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_clouds; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long mine = i < n_clouds ? (unsigned long long)cnt[i] : 0ull;
    unsigned long long v = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, v, o);
      if (lane >= o) v += y;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    unsigned long long before = s_carry;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (i < n_clouds) offs[i] = before + v - mine;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offs[n_clouds] = s_carry;
}

// selectWithinDistance with the final coefficients, per cloud, ascending: out[offs[c] + r] = index (within cloud c) of its
// r-th inlier.  One block per cloud walks the cloud in 1024-point steps with a block-wide scan per step.
template <int DOT>
__global__ void __launch_bounds__(256) batch_lists_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Z,
                                                          size_t n_per, size_t stride, const float4* __restrict__ planes,
                                                          const int32_t* __restrict__ best, float t,
                                                          const unsigned long long* __restrict__ offs, size_t cap, int32_t* __restrict__ out) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  pdl_wait();
  const size_t c = blockIdx.x;
  if (best[c] < 0) return;
  const float4 pl = planes[c];
  const unsigned long long o0 = offs[c];
  if (o0 + (offs[c + 1] - o0) > cap) return;  // the caller's buffer cannot hold this cloud's list: reported by the host
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const float* x = X + c * stride;
  const float* y = Y + c * stride;
  const float* z = Z + c * stride;
  for (size_t p0 = 0; p0 < n_per; p0 += 1024) {
    const size_t i0 = p0 + 4 * threadIdx.x;  // the planes are NaN-padded to the stride: no bounds check on the loads
    const float4 x4 = *reinterpret_cast<const float4*>(x + i0);
    const float4 y4 = *reinterpret_cast<const float4*>(y + i0);
    const float4 z4 = *reinterpret_cast<const float4*>(z + i0);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w}, zs[4] = {z4.x, z4.y, z4.z, z4.w};
    unsigned in = 0u;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float r;
      if (DOT == 1) r = __fmaf_rn(pl.x, xs[e], __fmaf_rn(pl.y, ys[e], __fmaf_rn(pl.z, zs[e], pl.w)));
      else r = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, xs[e]), __fmul_rn(pl.z, zs[e])), __fadd_rn(__fmul_rn(pl.y, ys[e]), pl.w));
      if (i0 + e < n_per && fabsf(r) < t) in |= 1u << e;
    }
    const int k = __popc(in);
    int v = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int yv = __shfl_up_sync(0xFFFFFFFFu, v, o);
      if (lane >= o) v += yv;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    int before = s_base;
    int total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) before += s_warp[w];
      total += s_warp[w];
    }
    int rnk = before + v - k;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (in & (1u << e)) out[o0 + (unsigned long long)(rnk++)] = (int32_t)(i0 + e);
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
}

// countWithinDistance of one plane per cloud (the final selection's size): block per cloud, HBM-bound.  (K2 with one
// hypothesis per cloud would leave 31 of 32 lanes of its hypothesis-per-lane layout empty.)
template <int DOT>
__global__ void __launch_bounds__(256) batch_count_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Z,
                                                          size_t n_per, size_t stride, const float4* __restrict__ planes,
                                                          const int32_t* __restrict__ best, float t, int32_t* __restrict__ cnt) {
  __shared__ int s_warp[8];
  pdl_wait();
  const size_t c = blockIdx.x;
  if (best[c] < 0) {
    if (threadIdx.x == 0) cnt[c] = 0;
    return;
  }
  const float4 pl = planes[c];
  const float* x = X + c * stride;
  const float* y = Y + c * stride;
  const float* z = Z + c * stride;
  int k = 0;
  for (size_t i0 = 4 * (size_t)threadIdx.x; i0 < n_per; i0 += 1024) {
    const float4 x4 = *reinterpret_cast<const float4*>(x + i0);
    const float4 y4 = *reinterpret_cast<const float4*>(y + i0);
    const float4 z4 = *reinterpret_cast<const float4*>(z + i0);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w}, zs[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float r;
      if (DOT == 1) r = __fmaf_rn(pl.x, xs[e], __fmaf_rn(pl.y, ys[e], __fmaf_rn(pl.z, zs[e], pl.w)));
      else r = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, xs[e]), __fmul_rn(pl.z, zs[e])), __fadd_rn(__fmul_rn(pl.y, ys[e]), pl.w));
      k += (i0 + e < n_per && fabsf(r) < t) ? 1 : 0;
    }
  }
  k = __reduce_add_sync(0xFFFFFFFFu, k);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < 8; ++w) total += s_warp[w];
    cnt[c] = total;
  }
}

void launch_batch_count(CloudView clouds, size_t n_per, size_t stride, int n_clouds, const float4* planes, const int32_t* best, float t,
                        int dot_order, int32_t* cnt, cudaStream_t s) {
  if (dot_order == 1) launch_chained(batch_count_kernel<1>, dim3(n_clouds), dim3(256), 0, s, clouds.x, clouds.y, clouds.z, n_per, stride, planes, best, t, cnt);
  else launch_chained(batch_count_kernel<0>, dim3(n_clouds), dim3(256), 0, s, clouds.x, clouds.y, clouds.z, n_per, stride, planes, best, t, cnt);
}

// K1a + K1b for a batch of equal-sized clouds that all draw the same K triples: thread (k, cloud) gathers its three sample
// points (kept: the refit takes its pivot from there), forms the model, and clears the count the scoring launch
// accumulates into; thread (0, 0) clears the batch's bad-sample flag.
__global__ void __launch_bounds__(128) batch_gather_models_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                                  const float* __restrict__ Z, size_t n_per, size_t stride,
                                                                  const int32_t* __restrict__ triples, int K, int4* __restrict__ sample_pts,
                                                                  float4* __restrict__ hyps, int32_t* __restrict__ good,
                                                                  int32_t* __restrict__ counts, int* __restrict__ flag) {
  pdl_wait();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t c = blockIdx.y;
  if (k == 0 && c == 0) *flag = 0;
  if (k >= K) return;
  const size_t o = c * (size_t)K + (size_t)k;
  int4 q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const long long j = (long long)triples[3 * k + i];
    q[i] = make_int4(0, 0, 0, 0);
    if (j >= 0 && j < (long long)n_per) {
      const size_t p = c * stride + (size_t)j;
      q[i] = make_int4(__float_as_int(X[p]), __float_as_int(Y[p]), __float_as_int(Z[p]), 0x3F800000);
    }
    sample_pts[3 * o + i] = q[i];
  }
  float4 h;
  const bool ok = model_from_sample(q[0], q[1], q[2], &h);
  hyps[o] = h;
  good[o] = ok ? 1 : 0;
  counts[o] = 0;
}

void launch_batch_gather_models(CloudView clouds, size_t n_per, size_t stride, int n_clouds, const int32_t* triples, int K, int4* sample_pts,
                                float4* hyps, int32_t* good, int32_t* counts, int* flag, cudaStream_t s) {
  launch_chained(batch_gather_models_kernel, dim3((K + 127) / 128, n_clouds), dim3(128), 0, s, clouds.x, clouds.y, clouds.z, n_per, stride, triples, K,
                 sample_pts, hyps, good, counts, flag);
}

void launch_batch_replay(const int32_t* counts, const int32_t* good, int K, int n_clouds, int32_t* best, int32_t* best_count, int* any_bad,
                         RefitOut* refit_to_clear, cudaStream_t s) {
  launch_chained(batch_replay_kernel, dim3(n_clouds), dim3(128), 0, s, counts, good, K, best, best_count, any_bad, refit_to_clear);
}

void launch_batch_finish(const float4* hyps, int K, const int32_t* best, const RefitOut* refit, const int32_t* scale_exp, int optimize,
                         int n_clouds, float4* raw, float4* refined, cudaStream_t s) {
  launch_chained(batch_finish_kernel, dim3((n_clouds + 63) / 64), dim3(64), 0, s, hyps, K, best, refit, scale_exp, optimize, n_clouds, raw, refined);
}

void launch_batch_lists(CloudView clouds, size_t n_per, size_t stride, int n_clouds, const float4* planes, const int32_t* best, float t,
                        int dot_order, const int32_t* cnt, unsigned long long* offs, size_t cap, int32_t* out, cudaStream_t s) {
  launch_chained(batch_offsets_kernel, dim3(1), dim3(1024), 0, s, cnt, n_clouds, offs);
  if (out == nullptr) return;
  if (dot_order == 1)
    launch_chained(batch_lists_kernel<1>, dim3(n_clouds), dim3(256), 0, s, clouds.x, clouds.y, clouds.z, n_per, stride, planes, best, t, offs, cap, out);
  else
    launch_chained(batch_lists_kernel<0>, dim3(n_clouds), dim3(256), 0, s, clouds.x, clouds.y, clouds.z, n_per, stride, planes, best, t, offs, cap, out);
}

}  // namespace pr
