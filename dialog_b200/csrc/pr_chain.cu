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

#include "pr_chain_dev.cuh"
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

// ---- round preparation: every buffer the round's kernels accumulate into, cleared by one launch ------------------------
__global__ void __launch_bounds__(256) round_prep_kernel(const RoundState* __restrict__ st, unsigned long long* __restrict__ table, size_t table_slots,
                                                         uint32_t* __restrict__ coll_count, int32_t* __restrict__ counts, int K,
                                                         RefitOut* __restrict__ refit, unsigned long long* __restrict__ scratch,
                                                         size_t scratch_words, unsigned* __restrict__ tickets) {
  if (st->stop) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = i0; i < table_slots; i += stride) table[i] = kDrawEmptySlot;
  for (size_t i = i0; i < scratch_words; i += stride) scratch[i] = 0ull;
  for (size_t i = i0; i < (size_t)K; i += stride) counts[i] = 0;
  if (i0 < sizeof(RefitOut) / sizeof(long long)) reinterpret_cast<long long*>(refit)[i0] = 0;
  if (i0 == 0) {
    *coll_count = 0u;
    tickets[0] = 0u;
  }
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
                                                                       const uint32_t* __restrict__ coll_count, RoundRecord* rec) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint32_t* s_sorted = reinterpret_cast<uint32_t*>(s_raw);
  int32_t* s_pre = reinterpret_cast<int32_t*>(s_raw + (size_t)kDrawCollCap * 4);
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(s_raw + (size_t)kDrawCollCap * 20);
  int32_t* s_vals = reinterpret_cast<int32_t*>(s_raw + (size_t)kDrawCollCap * 20 + (size_t)kResolveMapSlots * 4);
  __shared__ int s_distinct;
  if (st->stop) return;
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
  if (threadIdx.x != 0) return;
  if (s_distinct > kDrawMaxCollisions) {
    st->stop = 2;
    rec->stop = 2;
    __threadfence_system();
    rec->ran = 1;
    return;
  }
  SmemFetch fetch{s_sorted, s_pre};
  draw_resolve(s_sorted, (int)n_c, v, fetch, s_keys, s_vals, (uint32_t)(kResolveMapSlots - 1));
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
    rec->stop = 2;
    __threadfence_system();
    rec->ran = 1;
    return;
  }
  st->best = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
  st->best_count = (int)(best >> 32);
}

// ---- refined plane + stop rule as their own kernels (sharded clouds: an exchange sits between them and K3 / K5; on one
// GPU they run in the last block of refit_kernel / compact_kernel instead) ---------------------------------------------
__global__ void finish_kernel(RoundState* st, const float4* __restrict__ hyps, const int32_t* __restrict__ triples,
                              const RefitOut* __restrict__ refit, int optimize, int scale_exp, int n_draws, RoundRecord* rec) {
  if (threadIdx.x != 0 || st->stop) return;
  long long m[16];
  for (int i = 0; i < 16; ++i) m[i] = refit->m[i];
  const float pivot[3] = {refit->pivot[0], refit->pivot[1], refit->pivot[2]};
  chain_finish(st, hyps, triples, m, pivot, optimize, scale_exp, n_draws, rec);
}

__global__ void advance_kernel(RoundState* st, const long long* __restrict__ totals, int n_ranks, int rank, int min_plane,
                               RoundRecord* rec) {
  if (threadIdx.x != 0 || st->stop) return;
  chain_advance(st, totals, n_ranks, rank, min_plane, rec);
}

void launch_round_prep(const RoundState* st, unsigned long long* table, size_t table_slots, uint32_t* coll_count, int32_t* counts, int K,
                       RefitOut* refit, void* scratch, size_t scratch_bytes, unsigned* tickets, int num_sms, cudaStream_t s) {
  round_prep_kernel<<<num_sms * 2, 256, 0, s>>>(st, table, table_slots, coll_count, counts, K, refit,
                                               reinterpret_cast<unsigned long long*>(scratch), scratch_bytes / 8, tickets);
}

void launch_draw(const uint32_t* rnd, int n_draws, RoundState* st, int32_t* triples, unsigned long long* table, size_t table_slots,
                 uint32_t* coll, uint32_t* coll_count, RoundRecord* rec, cudaStream_t s) {
  const int n_ops = 3 * n_draws;
  int blocks = (n_ops + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  draw_scatter_kernel<<<blocks, 256, 0, s>>>(rnd, n_ops, st, triples, table, (uint32_t)(table_slots - 1), coll, coll_count);
  cudaFuncSetAttribute(draw_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResolveSmemBytes);  // per device: cheap, unconditional
  draw_resolve_kernel<<<1, kResolveThreads, kResolveSmemBytes, s>>>(st, triples, coll, coll_count, rec);
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
