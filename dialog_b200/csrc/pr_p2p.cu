// pr_p2p.cu — the per-round exchanges of the point-sharded path as peer-memory kernels over NVLink.
//
// Every exchange of a peel round is tiny (sample points 12 B x 3K, counts 4 B x K, 16 moments, 2 totals), so
// its cost is latency, not bandwidth.  Instead of one NCCL collective each (~40 us at 8 ranks), every rank
// owns a "mailbox" in its HBM that all peers map through CUDA IPC; a producer kernel stores its
// contribution straight into every peer's mailbox (plain st.global over NVLink), fences at system scope and
// raises a per-source flag (st.release.sys) carrying a monotonically increasing epoch; the consumer kernel
// that follows in the stream spins (bounded) on its own flags (ld.acquire.sys) and reduces the slots in rank
// order, which keeps integer sums bit-identical to the single-GPU result.  No flag is ever reset; a channel
// is reused only after another full exchange, which every peer can have passed only after it consumed the
// previous contents.
//
// The gather of the sample points is fused into the producer: the owner of a sampled point writes its
// coordinates directly into all peers' sample buffers.
#include "pr_kernels.h"

namespace pr {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Last CTA of the grid (ticket) publishes epoch into flag[rank] of every peer.
__device__ __forceinline__ void p2p_signal(const P2PView& v, size_t flag_off, unsigned long long epoch, unsigned* ticket) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0u;
      __threadfence_system();
      for (int r = 0; r < v.n_ranks; ++r)
        st_release_sys(reinterpret_cast<unsigned long long*>(v.peers[r] + flag_off) + v.rank, epoch);
    }
  }
}

// Thread 0 of the CTA waits until every source rank has published `epoch`; bounded so that a lost peer
// turns into an error code instead of a hung GPU.
__device__ __forceinline__ bool p2p_wait(const P2PView& v, size_t flag_off, unsigned long long epoch, unsigned* err) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(v.peers[v.rank] + flag_off);
    const long long t0 = clock64();
    int ok = 1;
    for (int r = 0; r < v.n_ranks && ok; ++r) {
      while (ld_acquire_sys(f + r) < epoch) {
        if (clock64() - t0 > 6000000000ll) {  // ~3 s
          atomicExch(err, 1u);
          ok = 0;
          break;
        }
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// ---- producers ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) p2p_push_kernel(P2PView v, const uint32_t* __restrict__ src, size_t n32, size_t dst_off,
                                                       size_t flag_off, unsigned long long epoch, unsigned* ticket) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < v.n_ranks; ++r) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(v.peers[r] + dst_off);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) dst[i] = src[i];
  }
  p2p_signal(v, flag_off, epoch, ticket);
}

// K1a fused with its exchange: the owner of sample s writes the point's bits into every peer's buffer.
__global__ void __launch_bounds__(256) p2p_gather_samples_kernel(P2PView v, const float* __restrict__ x, const float* __restrict__ y,
                                                                 const float* __restrict__ z, long long first, size_t n,
                                                                 const int32_t* __restrict__ triples, int n_samples,
                                                                 size_t sp_off, size_t flag_off, unsigned long long epoch,
                                                                 unsigned* ticket) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_samples; s += gridDim.x * blockDim.x) {
    const long long local = (long long)triples[s] - first;
    if (local >= 0 && local < (long long)n) {
      const int4 val = make_int4(__float_as_int(x[local]), __float_as_int(y[local]), __float_as_int(z[local]), 0x3F800000);
      for (int r = 0; r < v.n_ranks; ++r) reinterpret_cast<int4*>(v.peers[r] + sp_off)[s] = val;
    }
  }
  p2p_signal(v, flag_off, epoch, ticket);
}

// ---- consumers ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) p2p_wait_copy_kernel(P2PView v, size_t src_off, size_t flag_off, unsigned long long epoch,
                                                            uint32_t* __restrict__ dst, size_t n32, unsigned* err) {
  if (!p2p_wait(v, flag_off, epoch, err)) return;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(v.peers[v.rank] + src_off);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) dst[i] = __ldcg(src + i);
}

template <typename T>
__global__ void __launch_bounds__(256) p2p_wait_sum_kernel(P2PView v, size_t slot_off, size_t slot_stride, size_t flag_off,
                                                           unsigned long long epoch, T* __restrict__ dst, size_t n, unsigned* err) {
  if (!p2p_wait(v, flag_off, epoch, err)) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T acc = 0;
    for (int r = 0; r < v.n_ranks; ++r)  // fixed rank order
      acc += __ldcg(reinterpret_cast<const T*>(v.peers[v.rank] + slot_off + (size_t)r * slot_stride) + i);
    dst[i] = acc;
  }
}

static unsigned grid_for(size_t n) {
  size_t b = (n + 1023) / 1024;
  if (b < 1) b = 1;
  if (b > 32) b = 32;
  return (unsigned)b;
}

void launch_p2p_push(const P2PView& v, const void* src, size_t bytes, size_t dst_off, size_t flag_off, unsigned long long epoch,
                     unsigned* ticket, cudaStream_t s) {
  const size_t n32 = (bytes + 3) / 4;
  p2p_push_kernel<<<grid_for(n32), 256, 0, s>>>(v, reinterpret_cast<const uint32_t*>(src), n32, dst_off, flag_off, epoch, ticket);
}

void launch_p2p_gather_samples(const P2PView& v, CloudView cloud, long long first, size_t n, const int32_t* triples, int n_samples,
                               size_t sp_off, size_t flag_off, unsigned long long epoch, unsigned* ticket, cudaStream_t s) {
  p2p_gather_samples_kernel<<<grid_for((size_t)n_samples), 256, 0, s>>>(v, cloud.x, cloud.y, cloud.z, first, n, triples, n_samples,
                                                                        sp_off, flag_off, epoch, ticket);
}

void launch_p2p_wait_copy(const P2PView& v, size_t src_off, size_t flag_off, unsigned long long epoch, void* dst, size_t bytes,
                          unsigned* err, cudaStream_t s) {
  const size_t n32 = (bytes + 3) / 4;
  p2p_wait_copy_kernel<<<grid_for(n32), 256, 0, s>>>(v, src_off, flag_off, epoch, reinterpret_cast<uint32_t*>(dst), n32, err);
}

void launch_p2p_wait_sum_i32(const P2PView& v, size_t slot_off, size_t slot_stride, size_t flag_off, unsigned long long epoch,
                             int32_t* dst, size_t n, unsigned* err, cudaStream_t s) {
  p2p_wait_sum_kernel<int32_t><<<grid_for(n), 256, 0, s>>>(v, slot_off, slot_stride, flag_off, epoch, dst, n, err);
}

void launch_p2p_wait_sum_i64(const P2PView& v, size_t slot_off, size_t slot_stride, size_t flag_off, unsigned long long epoch,
                             long long* dst, size_t n, unsigned* err, cudaStream_t s) {
  p2p_wait_sum_kernel<long long><<<1, 256, 0, s>>>(v, slot_off, slot_stride, flag_off, epoch, dst, n, err);
}

}  // namespace pr
