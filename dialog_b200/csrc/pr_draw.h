// pr_draw.h — PCL's sample stream (SampleConsensusModel::drawIndexSample, pcl/sample_consensus/sac_model.h, the
// sampler behind the reference's only SAC call, Dialog/SimplifyVerticesSize.cpp:64-67) evaluated in parallel.
//
// The sequential definition: shuffled[] starts as the identity over the N indices of the round's cloud; draw t runs
// op s = 3t + i for i = 0, 1, 2:   swap(shuffled[i], shuffled[i + rnd_s % (N - i)])   and returns shuffled[0..3).
// rnd_s is the s-th value of mt19937(seed) >> 1 and does not depend on N, so the stream is generated once per seed.
//
// Write a_s = s % 3, b_s = a_s + rnd_s % (N - a_s), v_s = content of position b_s just before op s.  Op s leaves v_s
// in position a_s, and no later op of the same draw touches position a_s again (op i only touches positions >= i),
// so draw t returns (v_3t, v_3t+1, v_3t+2).  Position b_s holds b_s itself unless an earlier op touched it, which is
// rare (about (3K)^2 / 2N ops of a round collide): every op is therefore evaluated independently (v_s = b_s), the few
// ops whose position was picked by another op or lies in the head {0, 1, 2} are collected, sorted by s and replayed
// sequentially with the exact swap semantics.  draw_resolve below is that replay; the kernels in pr_kernels.cu and the
// host emulation behind plane_ransac_host_draw_triples_parallel (CPU tests against pr::IndexSampler) share it.
#pragma once

#include <stdint.h>

#include "pr_math.h"  // PR_HD

namespace pr {

constexpr uint32_t kDrawNoOp = 0xFFFFFFFFu;
// Collected ops the sequential replay accepts (more: the round falls back to the host sampler).
constexpr int kDrawMaxCollisions = 2048;
// Capacity of the collected list (an op can be listed more than once).
constexpr int kDrawCollCap = 3072;

PR_HD uint32_t draw_position(uint32_t s, uint32_t rnd, uint32_t n_points) {
  const uint32_t a = s % 3u;
  return a + rnd % (n_points - a);
}

PR_HD uint32_t draw_hash(uint32_t q) { return (q * 2654435761u) >> 5; }

// Sequential replay of the collected ops (ascending s, duplicates allowed).  v[s] holds b_s on entry for every op and
// the exact v_s on return.  The replay of op s reads v at s and at up to three earlier ops (s - 1, s - 2, s - 3: the last
// swaps of the head positions); fetch(i, idx) supplies those values as they were BEFORE the replay (i = position in
// ops_sorted; the kernel prefetches them into shared memory in parallel, the host reads v), and the three most
// recently replayed ops are kept in registers because their values have changed since.  map_*: open-addressing
// scratch with map_mask + 1 >= 2 * n_ops slots, keys preset to kDrawNoOp.
template <class Fetch>
PR_HD void draw_resolve(const uint32_t* ops_sorted, int n_ops, int32_t* v, Fetch fetch, uint32_t* map_keys, int32_t* map_vals,
                        uint32_t map_mask) {
  long long hw_time[3] = {-1, -1, -1};  // last swap of ANOTHER head position with head position p, and what it left there
  int32_t hw_val[3] = {0, 0, 0};
  long long done_s[3] = {-1, -1, -1};   // the three most recently replayed ops
  int32_t done_v[3] = {0, 0, 0};
  uint32_t prev = kDrawNoOp;
  for (int i = 0; i < n_ops; ++i) {
    const uint32_t s = ops_sorted[i];
    if (s == prev) continue;
    prev = s;
    const uint32_t a = s % 3u;
    const uint32_t q = (uint32_t)fetch(i, s);  // b_s: every op is replayed once
    // content of head position p just before op s: the later of the last op that had p as its first operand (every
    // third op; it left its v there) and the last recorded swap of another head position with p
    int32_t head[3];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (uint32_t p = 0; p < 3u; ++p) {
      const uint32_t d = (s % 3u + 3u - p) % 3u;
      const long long ua = (long long)s - (d == 0 ? 3 : (long long)d);  // negative: no such op yet
      int32_t val = (int32_t)p;
      if (hw_time[p] >= 0 && hw_time[p] > ua) {
        val = hw_val[p];
      } else if (ua >= 0) {
        val = fetch(i, (uint32_t)ua);
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int k = 0; k < 3; ++k)
          if (done_s[k] == ua) val = done_v[k];
      }
      head[p] = val;
    }
    const int32_t w = head[a];
    int32_t val;
    if (q < 3u) {
      val = head[q];
      if (q != a) {
        hw_time[q] = (long long)s;
        hw_val[q] = w;
      }
    } else {
      uint32_t h = draw_hash(q) & map_mask;
      while (map_keys[h] != kDrawNoOp && map_keys[h] != q) h = (h + 1) & map_mask;
      val = map_keys[h] == q ? map_vals[h] : (int32_t)q;
      map_keys[h] = q;
      map_vals[h] = w;
    }
    v[s] = val;
    done_s[2] = done_s[1]; done_v[2] = done_v[1];
    done_s[1] = done_s[0]; done_v[1] = done_v[0];
    done_s[0] = (long long)s; done_v[0] = val;
  }
}

// The common case needs no sequential replay at all.  While no listed op lies in the head (b < 3) the head position
// a = s % 3 is only ever touched by the ops s, s - 3, s - 6, ..., so op s finds there what op s - 3 left: v[s - 3] (or a
// itself for s < 3) — and leaves it in position b_s.  Hence the first op of a position group reads q itself and every
// later one reads what its predecessor s' in the group left there: the FINAL value of op s' - 3.  That op is either not
// listed (its value is its b) or listed, and then the same rule gives its value: a chain of strictly decreasing op
// indices that only reads the values the parallel phase wrote, so every entry is evaluated by its own thread.
// draw_independent_ok tests the condition for entry i, draw_resolve_independent gives entry i's value.
template <class Fetch>
PR_HD bool draw_independent_ok(const uint32_t* ops_sorted, int i, Fetch fetch) {
  return (uint32_t)fetch(i, ops_sorted[i]) >= 3u;
}

template <class Fetch>
PR_HD int32_t draw_resolve_independent(const uint32_t* ops_sorted, int i, Fetch fetch) {
  for (;;) {
    const uint32_t s = ops_sorted[i];
    const uint32_t q = (uint32_t)fetch(i, s);
    int j = i - 1;
    for (; j >= 0; --j) {
      const uint32_t sp = ops_sorted[j];
      if (sp != s && (uint32_t)fetch(j, sp) == q) break;  // (sp == s: a duplicate entry of this op)
    }
    if (j < 0) return (int32_t)q;
    const uint32_t sp = ops_sorted[j];
    if (sp < 3u) return (int32_t)(sp % 3u);
    const uint32_t u = sp - 3u;
    int lo = 0, hi = j;  // first entry >= u among the entries before j
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ops_sorted[mid] < u) lo = mid + 1;
      else hi = mid;
    }
    if (lo >= j || ops_sorted[lo] != u) return fetch(j, u);  // not listed: its b
    i = lo;
  }
}

// Parallel phase, one call per op: v[s] = b_s; ops in the head and ops whose position another op also picked are
// appended to coll (the first op of a position is appended by whoever finds it there; duplicates are fine).  The
// counter keeps counting past coll_cap so that the overflow is visible.  A: atomics policy (device / host emulation).
// The table is never cleared between rounds: a slot is position << 33 | op << 16 | epoch and counts as free unless its
// epoch is the current one (epochs run 1 .. 65535; the owner zeroes the table once and whenever the epoch wraps).
PR_HD unsigned long long draw_slot(uint32_t q, uint32_t s, uint32_t epoch) {
  return ((unsigned long long)q << 33) | ((unsigned long long)s << 16) | (unsigned long long)epoch;
}

template <class A>
PR_HD void draw_scatter(uint32_t s, uint32_t rnd, uint32_t n_points, int32_t* v, unsigned long long* table, uint32_t table_mask,
                        uint32_t epoch, uint32_t* coll, uint32_t* coll_count, uint32_t coll_cap) {
  const uint32_t q = draw_position(s, rnd, n_points);
  v[s] = (int32_t)q;
  uint32_t first_other = kDrawNoOp;
  bool collide = q < 3u;
  if (!collide) {
    const unsigned long long mine = draw_slot(q, s, epoch);
    uint32_t h = draw_hash(q) & table_mask;
    unsigned long long cur = A::load(&table[h]);
    for (;;) {
      if ((uint32_t)(cur & 0xFFFFull) != epoch) {  // free: left over from an earlier round
        const unsigned long long old = A::cas(&table[h], cur, mine);
        if (old == cur) break;                      // claimed
        cur = old;                                  // somebody was faster: look at what is there now
        continue;
      }
      if ((uint32_t)(cur >> 33) == q) {
        collide = true;
        first_other = (uint32_t)(cur >> 16) & 0x1FFFFu;
        break;
      }
      h = (h + 1) & table_mask;
      cur = A::load(&table[h]);
    }
  }
  if (collide) {
    const uint32_t n_new = first_other != kDrawNoOp ? 2u : 1u;
    const uint32_t at = A::add(coll_count, n_new);
    if (at < coll_cap) coll[at] = s;
    if (n_new == 2u && at + 1 < coll_cap) coll[at + 1] = first_other;
  }
}

}  // namespace pr
