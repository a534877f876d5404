// pr_chain_dev.cuh — device functions of the host-free peel loop shared by its own kernels (pr_chain.cu) and by the
// tails of K3 / K5 (pr_kernels.cu), which run them in their last block when the cloud is on one GPU (no exchange
// between the moments and the plane, or between the peel and the stop rule), saving two launches per round.
#pragma once

#include "pr_kernels.h"
#include "pr_math.h"

namespace pr {

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
  __threadfence_system();  // the record may live in mapped host memory: everything above lands before `ran`
  rec->ran = 1;
}

}  // namespace pr
