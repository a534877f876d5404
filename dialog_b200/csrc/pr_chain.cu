// pr_chain.cu — the small kernels that take the host out of the peel loop (score-all mode).
//
// One round of pcl::SACSegmentation::segment + ExtractIndices (the reference's peel at Dialog/PlaneDetect.h:1560-1566)
// is a fixed kernel sequence whose sizes and decisions live in a RoundState in HBM:
//   draw_scatter / draw_resolve   PCL's index triples for the round's cloud size (pr_draw.h)
//   gather + models + score       K1, K2 (pr_kernels.cu), sized from RoundState on the device
//   replay_kernel                 RandomSampleConsensus::computeModel's decision over the K counts
//   refit + finish_kernel         K3 moments -> pcl::eigen33 closed form (pr_math.h) -> refined plane
//   compact + advance_kernel      K5 peel, minimum-plane-size rule, sizes of the next round
// so the host queues whole rounds ahead and only reads the per-round records (pr_api.cpp run_chain).
#include "pr_kernels.h"

#include "pr_draw.h"
#include "pr_math.h"

namespace pr {

namespace {
struct DevAtomics {
  static __device__ __forceinline__ unsigned long long cas(unsigned long long* p, unsigned long long expect, unsigned long long v) {
    return atomicCAS(p, expect, v);
  }
  static __device__ __forceinline__ uint32_t add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
};
}  // namespace

// ---- sampler, parallel phase: one thread per op ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) draw_scatter_kernel(const uint32_t* __restrict__ rnd, int n_ops, const RoundState* __restrict__ st,
                                                           int32_t* __restrict__ v, unsigned long long* table, uint32_t table_mask,
                                                           uint32_t* coll, uint32_t* coll_count) {
  if (st->stop || st->n_global < 3) return;
  const uint32_t n_points = (uint32_t)st->n_global;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_ops; s += gridDim.x * blockDim.x)
    draw_scatter<DevAtomics>((uint32_t)s, rnd[s], n_points, v, table, table_mask, coll, coll_count, (uint32_t)kDrawCollCap);
}

// ---- sampler, sequential phase: the colliding ops, sorted by op index, replayed by one thread ------------------------------
constexpr int kResolveThreads = 1024;
constexpr int kResolveMapSlots = 2 * kDrawMaxCollisions;

__global__ void __launch_bounds__(kResolveThreads) draw_resolve_kernel(RoundState* st, int32_t* v, const uint32_t* __restrict__ coll,
                                                                       const uint32_t* __restrict__ coll_count, RoundRecord* rec) {
  __shared__ uint32_t s_sorted[kDrawCollCap];
  __shared__ uint32_t s_keys[kResolveMapSlots];
  __shared__ int32_t s_vals[kResolveMapSlots];
  __shared__ int s_distinct;
  if (st->stop) return;
  if (st->n_global < 3) {  // getSamples: "Can not select 0 unique points out of N": segment() returns no model
    if (threadIdx.x == 0) {
      st->stop = 1;
      st->best = -1;
      rec->ran = 1;
      rec->stop = 1;
      rec->n_cloud = st->n_global;
      rec->n_local = st->n_local;
      rec->n_rem_local = st->n_local;
      rec->n_rem_global = st->n_global;
      rec->first_after = st->first;
      rec->inl_off = st->inl_off;
    }
    return;
  }
  const uint32_t n_c = *coll_count;
  if (n_c > (uint32_t)kDrawCollCap) {  // crowded round: the sequential host sampler takes it
    if (threadIdx.x == 0) {
      st->stop = 2;
      rec->ran = 1;
      rec->stop = 2;
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
  for (uint32_t i = threadIdx.x; i < n_c; i += kResolveThreads)
    if (i == 0 || s_sorted[i] != s_sorted[i - 1]) atomicAdd(&s_distinct, 1);
  __syncthreads();
  if (threadIdx.x != 0) return;
  if (s_distinct > kDrawMaxCollisions) {
    st->stop = 2;
    rec->ran = 1;
    rec->stop = 2;
    return;
  }
  draw_resolve(s_sorted, (int)n_c, v, s_keys, s_vals, (uint32_t)(kResolveMapSlots - 1));
}

// ---- RandomSampleConsensus::computeModel over the K counts of a score-all round ------------------------------------------
// With probability 1 the loop scores max_iterations + 1 good samples and keeps the first one with the largest count
// (strict '>' against n_best = -INT_MAX).  A bad sample among the K draws means the loop needs more draws than were
// scored: the round goes back to the host loop (stop = 2), which replays PCL's redraw rule.
constexpr int kReplayThreads = 1024;

__global__ void __launch_bounds__(kReplayThreads) replay_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ good, int K,
                                                                RoundState* st, RoundRecord* rec) {
  __shared__ unsigned long long s_best[kReplayThreads / 32];
  if (st->stop) return;
  unsigned long long best = 0ull;
  int all_good = 1;
  for (int j = threadIdx.x; j < K; j += kReplayThreads) {
    if (!good[j]) all_good = 0;
    const unsigned long long key = ((unsigned long long)(unsigned)counts[j] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)j);
    best = key > best ? key : best;
  }
  all_good = __syncthreads_and(all_good);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_down_sync(0xFFFFFFFFu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int w = 1; w < kReplayThreads / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
  if (!all_good) {
    st->stop = 2;
    rec->ran = 1;
    rec->stop = 2;
    return;
  }
  st->best = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
  st->best_count = (int)(best >> 32);
}

// ---- refined plane of the round (PCL optimizeModelCoefficients) + the round's record ----------------------------------
__global__ void finish_kernel(RoundState* st, const float4* __restrict__ hyps, const int32_t* __restrict__ triples,
                              const RefitOut* __restrict__ refit, int optimize, int scale_exp, int n_draws, RoundRecord* rec) {
  if (threadIdx.x != 0 || st->stop) return;
  const int best = st->best;
  const float4 raw = hyps[best];
  float refined[4] = {raw.x, raw.y, raw.z, raw.w};
  if (optimize) {
    long long m[16];
    for (int i = 0; i < 16; ++i) m[i] = refit->m[i];
    const float pivot[3] = {refit->pivot[0], refit->pivot[1], refit->pivot[2]};
    pm_plane_from_moments(m, pivot, scale_exp, refined);  // < 4 inliers: keeps the sample's model
  }
  for (int i = 0; i < 4; ++i) st->plane[i] = refined[i];
  rec->ok = 1;
  rec->best = best;
  rec->best_count = st->best_count;
  for (int i = 0; i < 3; ++i) rec->best_sample[i] = triples[3 * best + i];
  rec->n_draws = n_draws;
  rec->raw[0] = raw.x; rec->raw[1] = raw.y; rec->raw[2] = raw.z; rec->raw[3] = raw.w;
  for (int i = 0; i < 4; ++i) rec->refined[i] = refined[i];
}

// ---- the peel's stop rule and the next round's sizes -------------------------------------------------------------------
// totals: [0] points left / [1] inliers peeled on this rank; sharded: [2 + 2r], [3 + 2r] the same for every rank r.
__global__ void advance_kernel(RoundState* st, const long long* __restrict__ totals, int n_ranks, int rank, int min_plane,
                               RoundRecord* rec) {
  if (threadIdx.x != 0 || st->stop) return;
  const long long rem_local = totals[0], inl_local = totals[1];
  long long rem_global = rem_local, inl_global = inl_local, first_after = 0;
  if (n_ranks > 1) {
    rem_global = 0;
    inl_global = 0;
    for (int r = 0; r < n_ranks; ++r) {
      if (r == rank) first_after = rem_global;
      rem_global += totals[2 + 2 * r];
      inl_global += totals[3 + 2 * r];
    }
  }
  rec->ran = 1;
  rec->n_cloud = st->n_global;
  rec->n_local = st->n_local;
  rec->n_inl_local = inl_local;
  rec->n_rem_local = rem_local;
  rec->n_inl_global = inl_global;
  rec->n_rem_global = rem_global;
  rec->first_after = first_after;
  rec->inl_off = st->inl_off;
  const long long need = min_plane > 0 ? (long long)min_plane : 0ll;
  const bool accepted = !(inl_global == 0 || inl_global < need);
  rec->accepted = accepted ? 1 : 0;
  if (accepted) {
    st->inl_off += inl_local;
    st->n_local = rem_local;
    st->n_global = rem_global;
    st->first = first_after;
    st->round += 1;
  } else {
    st->stop = 1;
  }
  rec->stop = st->stop;
}

void launch_draw(const uint32_t* rnd, int n_draws, RoundState* st, int32_t* triples, unsigned long long* table, size_t table_slots,
                 uint32_t* coll, uint32_t* coll_count, RoundRecord* rec, cudaStream_t s) {
  const int n_ops = 3 * n_draws;
  cudaMemsetAsync(table, 0xFF, table_slots * sizeof(unsigned long long), s);
  cudaMemsetAsync(coll_count, 0, sizeof(uint32_t), s);
  int blocks = (n_ops + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  draw_scatter_kernel<<<blocks, 256, 0, s>>>(rnd, n_ops, st, triples, table, (uint32_t)(table_slots - 1), coll, coll_count);
  draw_resolve_kernel<<<1, kResolveThreads, 0, s>>>(st, triples, coll, coll_count, rec);
}

size_t draw_table_slots(int n_draws) {
  size_t cap = 1024;
  while (cap < 12 * (size_t)n_draws) cap <<= 1;
  return cap;
}

void launch_replay(const int32_t* counts, const int32_t* good, int K, RoundState* st, RoundRecord* rec, cudaStream_t s) {
  replay_kernel<<<1, kReplayThreads, 0, s>>>(counts, good, K, st, rec);
}

void launch_finish(RoundState* st, const float4* hyps, const int32_t* triples, const RefitOut* refit, int optimize, int scale_exp,
                   int n_draws, RoundRecord* rec, cudaStream_t s) {
  finish_kernel<<<1, 32, 0, s>>>(st, hyps, triples, refit, optimize, scale_exp, n_draws, rec);
}

void launch_advance(RoundState* st, const long long* totals, int n_ranks, int rank, int min_plane, RoundRecord* rec, cudaStream_t s) {
  advance_kernel<<<1, 32, 0, s>>>(st, totals, n_ranks, rank, min_plane, rec);
}

}  // namespace pr
