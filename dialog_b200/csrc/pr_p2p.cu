// pr_p2p.cu — the per-round exchanges of the point-sharded path as peer-memory kernels over NVLink.
//
// Every exchange of a peel round is tiny (sample points 12 B x 3K, counts 4 B x K, 16 moments, 2 totals), so
// its cost is latency, not bandwidth.  Instead of one NCCL collective each (~40 us at 8 ranks), every rank
// owns a "mailbox" in its HBM that all peers map through CUDA IPC; one kernel per exchange stores this rank's
// contribution straight into every peer's mailbox (plain st.global over NVLink), fences at system scope,
// raises a per-source flag (st.release.sys) carrying a monotonically increasing epoch (kept on the device), spins (bounded) on its
// own flags (ld.acquire.sys) and reduces the slots in rank order, which keeps integer sums bit-identical to
// the single-GPU result.  No flag is ever reset; every channel has two buffers used alternately (epoch parity),
// and a peer can only rewrite a buffer after this rank produced the exchange in between, which in stream order
// follows this rank's reads of that buffer.
//
// The gather of the sample points is fused into the producer: the owner of a sampled point writes its
// coordinates directly into all peers' sample buffers.
#include "pr_kernels.h"

#include <algorithm>
#include <cstdlib>

#include "pr_chain_dev.cuh"

namespace pr {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// ---- one kernel per exchange: contribute, signal, wait, reduce -------------------------------------------------
// A single CTA does the whole exchange (the payloads are a few KB to ~150 KB): every thread stores its share of
// this rank's contribution into every peer's mailbox, fences at system scope, the CTA synchronises, thread r raises
// this rank's flag in peer r's mailbox (st.release.sys) and then waits for rank r's flag in the local mailbox
// (ld.acquire.sys, bounded), the CTA synchronises again and reduces / copies the slots in rank order with L1-bypassing
// loads.  One launch instead of a producer + consumer pair or an NCCL collective.
constexpr int kP2PThreads = 1024;

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// wait_ns (optional): [0] += the longest any thread of this exchange spun on a peer's flag (ns), [1] += 1 — what a rank
// loses per exchange to the slowest peer plus the NVLink round trip (pr_profile.p2p_wait_ms).
// signal: whether this block raises the rank's flags (one block does: the only one, or the last of several to have its
// stores out); every block waits for all ranks' flags.
__device__ __forceinline__ bool p2p_signal_and_wait(const P2PView& v, size_t flag_off, unsigned long long epoch, unsigned* err,
                                                    unsigned long long* wait_ns, bool signal = true, bool fence = true) {
  __shared__ int s_bad;
  __shared__ unsigned long long s_wait;
  if (threadIdx.x == 0) {
    s_bad = 0;
    s_wait = 0ull;
  }
  if (fence) __threadfence_system();  // (a system-scope fence behind remote stores costs 2-3 us: exactly one per exchange)
  __syncthreads();
  if ((int)threadIdx.x < v.n_ranks) {
    const int r = threadIdx.x;
    if (signal) st_release_sys(reinterpret_cast<unsigned long long*>(v.peers[r] + flag_off) + v.rank, epoch);
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(v.peers[v.rank] + flag_off) + r;
    const long long t0 = clock64();
    const unsigned long long w0 = global_ns();
    while (ld_acquire_sys(f) < epoch) {
      if (clock64() - t0 > 40000000000ll) {  // ~20 s: a lost peer becomes an error code, not a hung GPU
        atomicExch(err, 1u);
        s_bad = 1;
        break;
      }
    }
    atomicMax(&s_wait, global_ns() - w0);
  }
  __syncthreads();
  if (threadIdx.x == 0 && wait_ns != nullptr) {
    wait_ns[0] += s_wait;
    wait_ns[1] += 1ull;
  }
  return s_bad == 0;
}

// Epochs live on the device (one counter per channel, identical on every rank because every rank runs or skips the same
// exchanges): every thread reads the counter on entry, thread 0 stores the new value once the exchange is complete.
// A kernel queued by the host-free peel loop returns at once when the loop has stopped (st->stop, the same on every
// rank), without consuming an epoch, so two exchanges that use the same buffer always have a completed one between them.
__device__ __forceinline__ bool p2p_enter(const RoundState* st, const unsigned long long* epoch_ctr, const unsigned* err,
                                          unsigned long long* epoch, int tail_kind) {
  if (st != nullptr) pdl_wait();  // queued rounds launch their kernels with programmatic dependent launch (pr_chain_dev.cuh)
  if (st != nullptr && st->stop) return false;
  if (tail_kind != P2PTail::kNone)
    chain_stamp(st, tail_kind == P2PTail::kModels ? kStampModels : tail_kind == P2PTail::kReplay ? kStampDecide
                    : tail_kind == P2PTail::kFinish ? kStampFinish : kStampAdvance);
  if (*reinterpret_cast<const volatile unsigned*>(err)) return false;  // an earlier exchange timed out: do not wait 20 s again
  *epoch = *epoch_ctr + 1ull;
  return true;
}

// dst[i] = sum over ranks (rank order) of every rank's src[i]; in place is fine (src is read before the barrier).
template <typename T>
__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_kernel(P2PView v, const T* src, size_t n, size_t slot_off, size_t buffer_bytes,
                                                                    size_t slot_stride, size_t flag_off, unsigned long long* epoch_ctr,
                                                                    T* dst, unsigned* err, RoundState* st, unsigned long long* wait_ns,
                                                                    P2PTail tail) {
  unsigned long long epoch;
  if (!p2p_enter(st, epoch_ctr, err, &epoch, tail.kind)) return;
  slot_off += (size_t)(epoch & 1ull) * buffer_bytes;
  for (int r = 0; r < v.n_ranks; ++r) {
    T* out = reinterpret_cast<T*>(v.peers[r] + slot_off + (size_t)v.rank * slot_stride);
    for (size_t i = threadIdx.x; i < n; i += kP2PThreads) out[i] = src[i];
  }
  if (!p2p_signal_and_wait(v, flag_off, epoch, err, wait_ns)) return;
  for (size_t i = threadIdx.x; i < n; i += kP2PThreads) {
    T acc = 0;
    for (int r = 0; r < v.n_ranks; ++r)
      acc += __ldcg(reinterpret_cast<const T*>(v.peers[v.rank] + slot_off + (size_t)r * slot_stride) + i);
    dst[i] = acc;
  }
  if (threadIdx.x == 0) *epoch_ctr = epoch;
  // host-free peel loop: the step that consumes the sums runs in this CTA instead of a launch of its own
  if (tail.kind == P2PTail::kReplay) {
    __syncthreads();  // the sums are in dst (global memory, written by this block)
    chain_replay_block(reinterpret_cast<const int32_t*>(dst), tail.good, (int)n, st, tail.rec);
  } else if (tail.kind == P2PTail::kFinish) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const RefitOut* ro = reinterpret_cast<const RefitOut*>(dst);
      long long m[16];
      for (int i = 0; i < 16; ++i) m[i] = ro->m[i];
      const float pivot[3] = {ro->pivot[0], ro->pivot[1], ro->pivot[2]};
      chain_finish(st, tail.hyps, tail.triples, m, pivot, tail.optimize, tail.scale_exp, tail.n_draws, tail.rec);
    }
  }
}

// dst[r * n + i] = rank r's src[i]
template <typename T>
__global__ void __launch_bounds__(kP2PThreads) p2p_allgather_kernel(P2PView v, const T* __restrict__ src, size_t n, size_t slot_off,
                                                                    size_t buffer_bytes, size_t slot_stride, size_t flag_off,
                                                                    unsigned long long* epoch_ctr, T* __restrict__ dst, unsigned* err,
                                                                    RoundState* st, unsigned long long* wait_ns, P2PTail tail) {
  unsigned long long epoch;
  if (!p2p_enter(st, epoch_ctr, err, &epoch, tail.kind)) return;
  slot_off += (size_t)(epoch & 1ull) * buffer_bytes;
  for (int r = 0; r < v.n_ranks; ++r) {
    T* out = reinterpret_cast<T*>(v.peers[r] + slot_off + (size_t)v.rank * slot_stride);
    for (size_t i = threadIdx.x; i < n; i += kP2PThreads) out[i] = src[i];
  }
  if (!p2p_signal_and_wait(v, flag_off, epoch, err, wait_ns)) return;
  for (size_t i = threadIdx.x; i < n * (size_t)v.n_ranks; i += kP2PThreads) {
    const size_t r = i / n, k = i - r * n;
    dst[i] = __ldcg(reinterpret_cast<const T*>(v.peers[v.rank] + slot_off + r * slot_stride) + k);
  }
  if (threadIdx.x == 0) *epoch_ctr = epoch;
  if (tail.kind == P2PTail::kAdvance) {  // dst - n = this rank's own (remaining, inliers) pair, dst = every rank's
    __syncthreads();
    if (threadIdx.x == 0) chain_advance(st, reinterpret_cast<const long long*>(dst) - n, v.n_ranks, v.rank, tail.min_plane, tail.rec);
  }
}

// K1a fused with its exchange: the owner of sample s writes the point's bits into every rank's sample buffer; after
// the flags every rank copies the complete buffer out of its own mailbox.
__global__ void __launch_bounds__(kP2PThreads) p2p_samples_kernel(P2PView v, const float* __restrict__ x, const float* __restrict__ y,
                                                                  const float* __restrict__ z, long long first, size_t n,
                                                                  const int32_t* __restrict__ triples, int n_samples, size_t sp_off,
                                                                  size_t buffer_bytes, size_t flag_off, unsigned long long* epoch_ctr,
                                                                  int4* __restrict__ dst, unsigned* err, const RoundState* st,
                                                                  unsigned long long* wait_ns, P2PTail tail) {
  unsigned long long epoch;
  if (!p2p_enter(st, epoch_ctr, err, &epoch, tail.kind)) return;
  if (st != nullptr) {
    first = st->first;
    n = (size_t)st->n_local;
  }
  sp_off += (size_t)(epoch & 1ull) * buffer_bytes;
  // four samples per thread and pass, so that the index loads, then the 12 gathers, are in flight together (one CTA:
  // a dependent load chain per sample would cost ~1 us each)
  for (int base = threadIdx.x; base < n_samples; base += 4 * kP2PThreads) {
    long long local[4];
    float px[4], py[4], pz[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s = base + j * kP2PThreads;
      local[j] = s < n_samples ? (long long)triples[s] - first : -1ll;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool mine = local[j] >= 0 && local[j] < (long long)n;
      px[j] = mine ? x[local[j]] : 0.0f;
      py[j] = mine ? y[local[j]] : 0.0f;
      pz[j] = mine ? z[local[j]] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (local[j] >= 0 && local[j] < (long long)n) {
        const int4 val = make_int4(__float_as_int(px[j]), __float_as_int(py[j]), __float_as_int(pz[j]), 0x3F800000);
        for (int r = 0; r < v.n_ranks; ++r) reinterpret_cast<int4*>(v.peers[r] + sp_off)[base + j * kP2PThreads] = val;
      }
    }
  }
  if (!p2p_signal_and_wait(v, flag_off, epoch, err, wait_ns)) return;
  const int4* in = reinterpret_cast<const int4*>(v.peers[v.rank] + sp_off);
#pragma unroll 4
  for (int s = threadIdx.x; s < n_samples; s += kP2PThreads) dst[s] = __ldcg(in + s);
  if (threadIdx.x == 0) *epoch_ctr = epoch;
  if (tail.kind == P2PTail::kModels) {  // K1b in the same CTA: the models of the n_samples / 3 gathered triples
#pragma unroll 2
    for (int k = threadIdx.x; k < n_samples / 3; k += kP2PThreads) {
      float4 h;
      const bool ok = model_from_sample(__ldcg(in + 3 * k), __ldcg(in + 3 * k + 1), __ldcg(in + 3 * k + 2), &h);
      tail.hyps_out[k] = h;
      tail.good_out[k] = ok ? 1 : 0;
    }
  }
}

// ---- the two larger exchanges (sample points, counts) over several blocks ---------------------------------------------
// One SM keeps too few remote stores in flight: 196 KB of sample points through a single block take ~20 us.  With G
// blocks every block stores its slice, fences at system scope and takes a ticket; the block that takes the last ticket
// knows all of this rank's stores are out and raises the flags; all G blocks then wait for every rank's flag and consume
// their slice.  A second ticket finds the block that finishes last: it advances the channel's epoch (no block can get
// there before every block of this rank has read the epoch on entry: the flags it waits for include this rank's own),
// resets both tickets and runs the consuming step.  G <= 16 blocks on an otherwise idle GPU are co-resident, so the
// waiting blocks cannot starve the one that signals.
// remote: the block's stores went to peers (system-scope fence before the ticket); otherwise to local memory only.
__device__ __forceinline__ bool p2p_block_is_last(unsigned* ticket, bool remote) {
  __shared__ int s_last;
  if (remote) __threadfence_system();
  else __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (s_last) __threadfence();  // the other blocks' tickets, and with them their stores, are ordered before what follows
  return s_last != 0;
}

constexpr int kP2PMcThreads = 256;
constexpr int kP2PSampleBlocks = 16;

__global__ void __launch_bounds__(kP2PMcThreads) p2p_samples_mc_kernel(P2PView v, const float* __restrict__ x, const float* __restrict__ y,
                                                                       const float* __restrict__ z, long long first, size_t n,
                                                                       const int32_t* __restrict__ triples, int n_samples, size_t sp_off,
                                                                       size_t buffer_bytes, size_t flag_off, unsigned long long* epoch_ctr,
                                                                       int4* __restrict__ dst, unsigned* err, const RoundState* st,
                                                                       unsigned long long* wait_ns, P2PTail tail, unsigned* tickets) {
  unsigned long long epoch;
  if (!p2p_enter(st, epoch_ctr, err, &epoch, tail.kind)) return;
  if (st != nullptr) {
    first = st->first;
    n = (size_t)st->n_local;
  }
  sp_off += (size_t)(epoch & 1ull) * buffer_bytes;
  const int gtid = blockIdx.x * kP2PMcThreads + threadIdx.x, gstride = gridDim.x * kP2PMcThreads;
  for (int base = gtid; base < n_samples; base += 4 * gstride) {
    long long local[4];
    float px[4], py[4], pz[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s = base + j * gstride;
      local[j] = s < n_samples ? (long long)triples[s] - first : -1ll;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool mine = local[j] >= 0 && local[j] < (long long)n;
      px[j] = mine ? x[local[j]] : 0.0f;
      py[j] = mine ? y[local[j]] : 0.0f;
      pz[j] = mine ? z[local[j]] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (local[j] >= 0 && local[j] < (long long)n) {
        const int4 val = make_int4(__float_as_int(px[j]), __float_as_int(py[j]), __float_as_int(pz[j]), 0x3F800000);
        for (int r = 0; r < v.n_ranks; ++r) reinterpret_cast<int4*>(v.peers[r] + sp_off)[base + j * gstride] = val;
      }
    }
  }
  const bool signal = p2p_block_is_last(&tickets[0], true);
  if (!p2p_signal_and_wait(v, flag_off, epoch, err, blockIdx.x == 0 ? wait_ns : nullptr, signal, false)) return;
  const int4* in = reinterpret_cast<const int4*>(v.peers[v.rank] + sp_off);
#pragma unroll 4
  for (int s = gtid; s < n_samples; s += gstride) dst[s] = __ldcg(in + s);
  if (tail.kind == P2PTail::kModels) {  // K1b: this block's share of the n_samples / 3 models
    for (int k = gtid; k < n_samples / 3; k += gstride) {
      float4 h;
      const bool ok = model_from_sample(__ldcg(in + 3 * k), __ldcg(in + 3 * k + 1), __ldcg(in + 3 * k + 2), &h);
      tail.hyps_out[k] = h;
      tail.good_out[k] = ok ? 1 : 0;
    }
  }
  if (p2p_block_is_last(&tickets[1], false) && threadIdx.x == 0) {
    *epoch_ctr = epoch;
    tickets[0] = 0u;
    tickets[1] = 0u;
  }
}

// counts: dst[i] = sum over ranks of src[i], G blocks of kP2PThreads; the block that finishes last runs computeModel's
// decision over the complete sums (tail.kind == kReplay).
__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_i32_mc_kernel(P2PView v, const int32_t* src, size_t n, size_t slot_off,
                                                                           size_t buffer_bytes, size_t slot_stride, size_t flag_off,
                                                                           unsigned long long* epoch_ctr, int32_t* dst, unsigned* err,
                                                                           RoundState* st, unsigned long long* wait_ns, P2PTail tail,
                                                                           unsigned* tickets) {
  unsigned long long epoch;
  if (!p2p_enter(st, epoch_ctr, err, &epoch, tail.kind)) return;
  slot_off += (size_t)(epoch & 1ull) * buffer_bytes;
  const size_t gtid = (size_t)blockIdx.x * kP2PThreads + threadIdx.x, gstride = (size_t)gridDim.x * kP2PThreads;
  for (size_t i = gtid; i < n; i += gstride) {
    const int32_t mine = src[i];
    for (int r = 0; r < v.n_ranks; ++r)
      reinterpret_cast<int32_t*>(v.peers[r] + slot_off + (size_t)v.rank * slot_stride)[i] = mine;
  }
  const bool signal = p2p_block_is_last(&tickets[0], true);
  if (!p2p_signal_and_wait(v, flag_off, epoch, err, blockIdx.x == 0 ? wait_ns : nullptr, signal, false)) return;
  for (size_t i = gtid; i < n; i += gstride) {  // the same slice this block read from src above: in place is fine
    int32_t acc = 0;
    for (int r = 0; r < v.n_ranks; ++r)
      acc += __ldcg(reinterpret_cast<const int32_t*>(v.peers[v.rank] + slot_off + (size_t)r * slot_stride) + i);
    dst[i] = acc;
  }
  if (!p2p_block_is_last(&tickets[1], false)) return;
  if (threadIdx.x == 0) {
    *epoch_ctr = epoch;
    tickets[0] = 0u;
    tickets[1] = 0u;
  }
  if (tail.kind == P2PTail::kReplay)  // every block's sums are in dst and visible (fence + ticket)
    chain_replay_block(reinterpret_cast<const int32_t*>(dst), tail.good, (int)n, st, tail.rec);
}

void launch_p2p_allreduce_i32(const P2PView& v, const int32_t* src, size_t n, size_t slot_off, size_t buffer_bytes, size_t slot_stride,
                              size_t flag_off, unsigned long long* epoch_ctr, int32_t* dst, unsigned* err, cudaStream_t s, RoundState* st,
                              unsigned long long* wait_ns, const P2PTail* tail, unsigned* tickets) {
  static const bool mc = [] { const char* e = getenv("PR_P2P_MC"); return !(e && atoi(e) == 0); }();
  const unsigned blocks = (unsigned)std::min<size_t>(8, (n + kP2PThreads - 1) / kP2PThreads);
  if (mc && tickets != nullptr && blocks > 1) {
    if (st != nullptr)
      launch_chained(p2p_allreduce_i32_mc_kernel, dim3(blocks), dim3(kP2PThreads), 0, s, v, src, n, slot_off, buffer_bytes, slot_stride, flag_off,
                     epoch_ctr, dst, err, st, wait_ns, tail ? *tail : P2PTail(), tickets);
    else
      p2p_allreduce_i32_mc_kernel<<<blocks, kP2PThreads, 0, s>>>(v, src, n, slot_off, buffer_bytes, slot_stride, flag_off, epoch_ctr, dst, err, st,
                                                                 wait_ns, tail ? *tail : P2PTail(), tickets);
    return;
  }
  if (st != nullptr)
    launch_chained(p2p_allreduce_kernel<int32_t>, dim3(1), dim3(kP2PThreads), 0, s, v, src, n, slot_off, buffer_bytes, slot_stride, flag_off, epoch_ctr,
                   dst, err, st, wait_ns, tail ? *tail : P2PTail());
  else
    p2p_allreduce_kernel<int32_t><<<1, kP2PThreads, 0, s>>>(v, src, n, slot_off, buffer_bytes, slot_stride, flag_off, epoch_ctr, dst, err, st,
                                                            wait_ns, tail ? *tail : P2PTail());
}

void launch_p2p_allreduce_i64(const P2PView& v, const long long* src, size_t n, size_t slot_off, size_t buffer_bytes, size_t slot_stride,
                              size_t flag_off, unsigned long long* epoch_ctr, long long* dst, unsigned* err, cudaStream_t s, RoundState* st,
                              unsigned long long* wait_ns, const P2PTail* tail) {
  if (st != nullptr)
    launch_chained(p2p_allreduce_kernel<long long>, dim3(1), dim3(kP2PThreads), 0, s, v, src, n, slot_off, buffer_bytes, slot_stride, flag_off,
                   epoch_ctr, dst, err, st, wait_ns, tail ? *tail : P2PTail());
  else
    p2p_allreduce_kernel<long long><<<1, kP2PThreads, 0, s>>>(v, src, n, slot_off, buffer_bytes, slot_stride, flag_off, epoch_ctr, dst, err, st,
                                                              wait_ns, tail ? *tail : P2PTail());
}

void launch_p2p_allgather_i64(const P2PView& v, const long long* src, size_t n, size_t slot_off, size_t buffer_bytes, size_t slot_stride,
                              size_t flag_off, unsigned long long* epoch_ctr, long long* dst, unsigned* err, cudaStream_t s, RoundState* st,
                              unsigned long long* wait_ns, const P2PTail* tail) {
  if (st != nullptr)
    launch_chained(p2p_allgather_kernel<long long>, dim3(1), dim3(kP2PThreads), 0, s, v, src, n, slot_off, buffer_bytes, slot_stride, flag_off,
                   epoch_ctr, dst, err, st, wait_ns, tail ? *tail : P2PTail());
  else
    p2p_allgather_kernel<long long><<<1, kP2PThreads, 0, s>>>(v, src, n, slot_off, buffer_bytes, slot_stride, flag_off, epoch_ctr, dst, err, st,
                                                              wait_ns, tail ? *tail : P2PTail());
}

void launch_p2p_samples(const P2PView& v, CloudView cloud, long long first, size_t n, const int32_t* triples, int n_samples,
                        size_t sp_off, size_t buffer_bytes, size_t flag_off, unsigned long long* epoch_ctr, int4* dst, unsigned* err,
                        cudaStream_t s, const RoundState* st, unsigned long long* wait_ns, const P2PTail* tail, unsigned* tickets) {
  static const bool mc = [] { const char* e = getenv("PR_P2P_MC"); return !(e && atoi(e) == 0); }();
  if (mc && tickets != nullptr && n_samples > 4 * kP2PMcThreads) {
    const unsigned blocks = (unsigned)std::min(kP2PSampleBlocks, (n_samples + 4 * kP2PMcThreads - 1) / (4 * kP2PMcThreads));
    if (st != nullptr)
      launch_chained(p2p_samples_mc_kernel, dim3(blocks), dim3(kP2PMcThreads), 0, s, v, cloud.x, cloud.y, cloud.z, first, n, triples, n_samples,
                     sp_off, buffer_bytes, flag_off, epoch_ctr, dst, err, st, wait_ns, tail ? *tail : P2PTail(), tickets);
    else
      p2p_samples_mc_kernel<<<blocks, kP2PMcThreads, 0, s>>>(v, cloud.x, cloud.y, cloud.z, first, n, triples, n_samples, sp_off, buffer_bytes,
                                                             flag_off, epoch_ctr, dst, err, st, wait_ns, tail ? *tail : P2PTail(), tickets);
    return;
  }
  if (st != nullptr)
    launch_chained(p2p_samples_kernel, dim3(1), dim3(kP2PThreads), 0, s, v, cloud.x, cloud.y, cloud.z, first, n, triples, n_samples, sp_off,
                   buffer_bytes, flag_off, epoch_ctr, dst, err, st, wait_ns, tail ? *tail : P2PTail());
  else
    p2p_samples_kernel<<<1, kP2PThreads, 0, s>>>(v, cloud.x, cloud.y, cloud.z, first, n, triples, n_samples, sp_off, buffer_bytes, flag_off,
                                                epoch_ctr, dst, err, st, wait_ns, tail ? *tail : P2PTail());
}

}  // namespace pr
