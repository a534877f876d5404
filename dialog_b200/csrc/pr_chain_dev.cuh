// pr_chain_dev.cuh — device functions of the host-free peel loop shared by its own kernels (pr_chain.cu) and by the
// tails of K3 / K5 (pr_kernels.cu), which run them in their last block when the cloud is on one GPU (no exchange
// between the moments and the plane, or between the peel and the stop rule), saving two launches per round.
#pragma once

#include <math_constants.h>

#include <cstdlib>
#include <tuple>
#include <utility>

#include "pr_kernels.h"
#include "pr_math.h"

namespace pr {

// ---- programmatic dependent launch for the kernels of a queued round -----------------------------------------------------
// Every kernel of the host-free loop depends on the one before it, so the ~2-3 us it takes to set up and schedule a grid
// would sit exposed between any two of them (about nine per round).  Launched with the programmatic-stream-serialization
// attribute, a kernel's grid is set up while its predecessor drains; the kernel then waits at griddepcontrol.wait (first
// statement) until the predecessor has completed and its writes are visible.  PR_PDL=0 launches them plainly.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ unsigned long long chain_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// call right after pdl_wait(): the predecessor of this kernel has just completed
__device__ __forceinline__ void chain_stamp(const RoundState* st, int slot) {
  if (st != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) const_cast<RoundState*>(st)->t[slot] = chain_now();
}

template <typename... P, size_t... I>
inline cudaError_t launch_chained_impl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, std::tuple<P...>& params,
                                       std::index_sequence<I...>) {
  static const bool pdl = [] { const char* e = getenv("PR_PDL"); return !(e && atoi(e) == 0); }();
  void* ptrs[] = {static_cast<void*>(&std::get<I>(params))...};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kernel), ptrs);
}

template <typename... P, typename... A>
inline cudaError_t launch_chained(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A... args) {
  std::tuple<P...> params(static_cast<P>(args)...);  // the kernel's own parameter types
  return launch_chained_impl(kernel, grid, block, smem, s, params, std::index_sequence_for<P...>{});
}

// PCL SampleConsensusModelPlane::isSampleGood + computeModelCoefficients (sac_model_plane.hpp), FP32,
// every operation rounded on its own; reductions in Eigen's SSE2 order (e0 + e2) + (e1 + e3).
__device__ __forceinline__ bool model_from_sample(int4 q0, int4 q1, int4 q2, float4* out) {
  float p0x = __int_as_float(q0.x), p0y = __int_as_float(q0.y), p0z = __int_as_float(q0.z);
  float ux = __fsub_rn(__int_as_float(q1.x), p0x), uy = __fsub_rn(__int_as_float(q1.y), p0y),
        uz = __fsub_rn(__int_as_float(q1.z), p0z);
  float vx = __fsub_rn(__int_as_float(q2.x), p0x), vy = __fsub_rn(__int_as_float(q2.y), p0y),
        vz = __fsub_rn(__int_as_float(q2.z), p0z);
  float r0 = __fdiv_rn(ux, vx), r1 = __fdiv_rn(uy, vy), r2 = __fdiv_rn(uz, vz);
  bool ok = (r0 != r1) || (r2 != r1);
  float4 h = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
  if (ok) {
    float nx = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy));
    float ny = __fsub_rn(__fmul_rn(uz, vx), __fmul_rn(ux, vz));
    float nz = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
    float sq = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(nz, nz)), __fadd_rn(__fmul_rn(ny, ny), 0.0f));
    float nrm = __fsqrt_rn(sq);
    nx = __fdiv_rn(nx, nrm);
    ny = __fdiv_rn(ny, nrm);
    nz = __fdiv_rn(nz, nrm);
    float dot = __fadd_rn(__fadd_rn(__fmul_rn(nx, p0x), __fmul_rn(nz, p0z)), __fadd_rn(__fmul_rn(ny, p0y), 0.0f));
    h = make_float4(nx, ny, nz, __fmul_rn(-1.0f, dot));
  }
  *out = h;
  return ok;
}

// RandomSampleConsensus::computeModel over the K counts of a score-all round, by one block of kChainBlock threads (all
// of them must call it).  With probability 1 the loop scores max_iterations + 1 good samples and keeps the first one
// with the largest count (strict '>' against n_best = -INT_MAX).  A bad sample among the K draws means the loop needs
// more draws than were scored: the round goes back to the host loop (stop = 2), which replays PCL's redraw rule.
constexpr int kChainBlock = 1024;
__device__ __forceinline__ void chain_replay_block(const int32_t* counts, const int32_t* __restrict__ good, int K, RoundState* st,
                                                   RoundRecord* rec) {
  __shared__ unsigned long long s_best[kChainBlock / 32];
  unsigned long long best = 0ull;
  int all_good = 1;
  for (int j = threadIdx.x; j < K; j += kChainBlock) {
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
  for (int w = 1; w < kChainBlock / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
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

// Refined plane of the round (PCL optimizeModelCoefficients: closed form from the summed moments, pr_math.h) + the
// round's record.  m: the 16 moments as this thread sees them (already summed over blocks / ranks).
__device__ __forceinline__ void chain_finish(RoundState* st, const float4* __restrict__ hyps, const int32_t* __restrict__ triples,
                                             const long long m[16], const float pivot[3], int optimize, int scale_exp, int n_draws,
                                             RoundRecord* rec) {
  const int best = st->best;
  const float4 raw = hyps[best];
  float refined[4] = {raw.x, raw.y, raw.z, raw.w};
  if (optimize) pm_plane_from_moments(m, pivot, scale_exp, refined);  // < 4 inliers: keeps the sample's model
  for (int i = 0; i < 4; ++i) st->plane[i] = refined[i];
  rec->ok = 1;
  rec->best = best;
  rec->best_count = st->best_count;
  for (int i = 0; i < 3; ++i) rec->best_sample[i] = triples[3 * best + i];
  rec->n_draws = n_draws;
  rec->raw[0] = raw.x; rec->raw[1] = raw.y; rec->raw[2] = raw.z; rec->raw[3] = raw.w;
  for (int i = 0; i < 4; ++i) rec->refined[i] = refined[i];
}

// The peel's stop rule and the next round's sizes.  totals: [0] points left / [1] inliers peeled on this rank;
// sharded: [2 + 2r], [3 + 2r] the same for every rank r.
__device__ __forceinline__ void chain_advance(RoundState* st, const long long* totals, int n_ranks, int rank, int min_plane,
                                              RoundRecord* rec) {
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
  for (int i = 0; i < kStampEnd; ++i) {
    rec->t[i] = st->t[i];
    st->t[i] = 0ull;
  }
  rec->t[kStampEnd] = chain_now();
  __threadfence_system();  // the record may live in mapped host memory: everything above lands before `ran`
  rec->ran = 1;
}

}  // namespace pr
