// Microbenchmark: issue-rate experiments for the hypothesis-scoring inner loop on sm_100a.
// Not part of the product; used to choose the kernel structure and to measure the FP32 FMA peak.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int TP = 1024;

// pure FFMA peak: 16 independent chains
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
  float r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b) {
  float2 r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = make_float2(threadIdx.x * 0.001f + i, i);
  float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = __ffma2_rn(r[i], A, B);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i].x + r[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_fset(float* out, int iters, float t) {
  float r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("set.lt.f32.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(t));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_iadd3(int* out, int iters, int a, int b) {
  int r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(r[i]) : "r"(a), "r"(r[(i + 1) & 15]));
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NI, bool PACKED>
__global__ void __launch_bounds__(256) k_mix(float* out, int iters, float a, float b, int c) {
  float2 r[12]; int q[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) { r[i] = make_float2(threadIdx.x * 0.001f + i, i); q[i] = threadIdx.x + i; }
  float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        if (PACKED) r[i] = __ffma2_rn(r[i], A, B);
        else { r[i].x = fmaf(r[i].x, a, b); r[i].y = fmaf(r[i].y, a, b); }
        if (i < NI) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(q[i]) : "r"(c), "r"(q[(i + 1) % 12]));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) s += r[i].x + r[i].y + q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NI, bool PACKED>
void run_mix(int nsm, float* dout) {
  int grid = nsm * 4, iters = 2048;
  float ms = time_ms([&] { k_mix<NI, PACKED><<<grid, 256>>>(dout, iters, 1.0001f, 0.5f, 3); });
  // cycles per (12 fma-pairs + NI iadd3) group per SMSP: warps per SMSP = grid*8/nsm/4 = 8
  double groups = (double)iters * 4 * 8;  // per SMSP
  printf("mix %s 12 fma-pairs + %2d IADD3: %.3f ms -> %.2f cycles per group per SMSP-warp-slot\n", PACKED ? "FFMA2" : "FFMA ", NI, ms, ms * 1e-3 * 1.965e9 / groups);
}
// MODE 0: scalar FFMA, int count     (hyps a,b,c,d scalar regs)
// MODE 1: FFMA2 (2 points per op, hyps duplicated), int count
// MODE 2: FFMA2, float count via FADD2
// MODE 3: FFMA2, mask-popc count (8 predicates -> bits -> popc)
// MODE 4: FFMA2, int count via 3-input add of two selects
template <int H, int MODE>
__global__ void __launch_bounds__(256, 2) k_score(const float4* __restrict__ hyps, int* __restrict__ out, int iters, float t) {
  __shared__ __align__(16) float sx[TP];
  __shared__ __align__(16) float sy[TP];
  __shared__ __align__(16) float sz[TP];
  for (int i = threadIdx.x; i < TP; i += blockDim.x) {
    sx[i] = (i * 37 % 101) * 0.03f; sy[i] = (i * 11 % 97) * 0.031f; sz[i] = (i * 7 % 89) * 0.029f;
  }
  __syncthreads();
  float a[H], b[H], c[H], d[H];
  int cnt[H];
  float2 fc[H];
  unsigned fcu[H];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    float4 h = hyps[(blockIdx.x * blockDim.x + threadIdx.x) * H + j];
    a[j] = h.x; b[j] = h.y; c[j] = h.z; d[j] = h.w; cnt[j] = 0; fc[j] = make_float2(0.f, 0.f); fcu[j] = 0;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int pb = 0; pb < TP; pb += 4 * 64) {
    if (MODE == 7 || MODE == 8) {
#pragma unroll
      for (int j = 0; j < H; ++j) { cnt[j] += ((fcu[j] >> 23) * 383u) & 511u; fcu[j] = 0; }
    }
#pragma unroll 2
    for (int p = pb; p < pb + 4 * 64; p += 4) {
      float4 X = *reinterpret_cast<const float4*>(&sx[p]);
      float4 Y = *reinterpret_cast<const float4*>(&sy[p]);
      float4 Z = *reinterpret_cast<const float4*>(&sz[p]);
      if (MODE == 8) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
          float r0 = fmaf(a[j], X.x, fmaf(b[j], Y.x, fmaf(c[j], Z.x, d[j])));
          float r1 = fmaf(a[j], X.y, fmaf(b[j], Y.y, fmaf(c[j], Z.y, d[j])));
          float r2 = fmaf(a[j], X.z, fmaf(b[j], Y.z, fmaf(c[j], Z.z, d[j])));
          float r3 = fmaf(a[j], X.w, fmaf(b[j], Y.w, fmaf(c[j], Z.w, d[j])));
          float f0, f1, f2, f3;
          asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f0) : "f"(fabsf(r0)), "f"(t));
          asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f1) : "f"(fabsf(r1)), "f"(t));
          asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f2) : "f"(fabsf(r2)), "f"(t));
          asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f3) : "f"(fabsf(r3)), "f"(t));
          unsigned acc = fcu[j];
          acc = acc + __float_as_uint(f0) + __float_as_uint(f1);
          acc = acc + __float_as_uint(f2) + __float_as_uint(f3);
          fcu[j] = acc;
        }
      } else if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
          float r0 = fmaf(a[j], X.x, fmaf(b[j], Y.x, fmaf(c[j], Z.x, d[j])));
          float r1 = fmaf(a[j], X.y, fmaf(b[j], Y.y, fmaf(c[j], Z.y, d[j])));
          float r2 = fmaf(a[j], X.z, fmaf(b[j], Y.z, fmaf(c[j], Z.z, d[j])));
          float r3 = fmaf(a[j], X.w, fmaf(b[j], Y.w, fmaf(c[j], Z.w, d[j])));
          cnt[j] += (fabsf(r0) < t); cnt[j] += (fabsf(r1) < t);
          cnt[j] += (fabsf(r2) < t); cnt[j] += (fabsf(r3) < t);
        }
      } else {
        float2 X0 = make_float2(X.x, X.y), X1 = make_float2(X.z, X.w);
        float2 Y0 = make_float2(Y.x, Y.y), Y1 = make_float2(Y.z, Y.w);
        float2 Z0 = make_float2(Z.x, Z.y), Z1 = make_float2(Z.z, Z.w);
#pragma unroll
        for (int j = 0; j < H; ++j) {
          float2 A = make_float2(a[j], a[j]), B = make_float2(b[j], b[j]);
          float2 C = make_float2(c[j], c[j]), D = make_float2(d[j], d[j]);
          float2 r0 = __ffma2_rn(A, X0, __ffma2_rn(B, Y0, __ffma2_rn(C, Z0, D)));
          float2 r1 = __ffma2_rn(A, X1, __ffma2_rn(B, Y1, __ffma2_rn(C, Z1, D)));
          if (MODE == 1) {
            cnt[j] += (fabsf(r0.x) < t); cnt[j] += (fabsf(r0.y) < t);
            cnt[j] += (fabsf(r1.x) < t); cnt[j] += (fabsf(r1.y) < t);
          } else if (MODE == 2) {
            float2 i0 = make_float2(fabsf(r0.x) < t ? 1.f : 0.f, fabsf(r0.y) < t ? 1.f : 0.f);
            float2 i1 = make_float2(fabsf(r1.x) < t ? 1.f : 0.f, fabsf(r1.y) < t ? 1.f : 0.f);
            fc[j] = __fadd2_rn(fc[j], i0); fc[j] = __fadd2_rn(fc[j], i1);
          } else if (MODE == 3) {
            unsigned m = (fabsf(r0.x) < t ? 1u : 0u) | (fabsf(r0.y) < t ? 2u : 0u) |
                         (fabsf(r1.x) < t ? 4u : 0u) | (fabsf(r1.y) < t ? 8u : 0u);
            cnt[j] += __popc(m);
          } else if (MODE == 5) {
            asm("{.reg .pred p; setp.lt.f32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(cnt[j]) : "f"(fabsf(r0.x)), "f"(t));
            asm("{.reg .pred p; setp.lt.f32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(cnt[j]) : "f"(fabsf(r0.y)), "f"(t));
            asm("{.reg .pred p; setp.lt.f32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(cnt[j]) : "f"(fabsf(r1.x)), "f"(t));
            asm("{.reg .pred p; setp.lt.f32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(cnt[j]) : "f"(fabsf(r1.y)), "f"(t));
          } else if (MODE == 6) {
            int m0, m1, m2, m3;
            asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(m0) : "f"(fabsf(r0.x)), "f"(t));
            asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(m1) : "f"(fabsf(r0.y)), "f"(t));
            asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(m2) : "f"(fabsf(r1.x)), "f"(t));
            asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(m3) : "f"(fabsf(r1.y)), "f"(t));
            cnt[j] = cnt[j] - m0 - m1; cnt[j] = cnt[j] - m2 - m3;
          } else if (MODE == 7) {
            // FSET.BF (1.0f / 0.0f bit patterns) summed two at a time by one IADD3; the count is
            // recovered mod 512 from (c * 0x3F800000) mod 2^32 and flushed before it can wrap.
            float f0, f1, f2, f3;
            asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f0) : "f"(fabsf(r0.x)), "f"(t));
            asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f1) : "f"(fabsf(r0.y)), "f"(t));
            asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f2) : "f"(fabsf(r1.x)), "f"(t));
            asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(f3) : "f"(fabsf(r1.y)), "f"(t));
            unsigned acc = (unsigned)fcu[j];
            acc = acc + __float_as_uint(f0) + __float_as_uint(f1);
            acc = acc + __float_as_uint(f2) + __float_as_uint(f3);
            fcu[j] = acc;
          } else if (MODE == 4) {
            int s0 = (fabsf(r0.x) < t) ? 1 : 0, s1 = (fabsf(r0.y) < t) ? 1 : 0;
            int s2 = (fabsf(r1.x) < t) ? 1 : 0, s3 = (fabsf(r1.y) < t) ? 1 : 0;
            cnt[j] = cnt[j] + s0 + s1; cnt[j] = cnt[j] + s2 + s3;
          }
        }
      }
    }
    }
  }
#pragma unroll
  for (int j = 0; j < H; ++j) {
    int v = cnt[j] + (int)(fc[j].x + fc[j].y) + (int)(((fcu[j] >> 23) * 383u) & 511u);
    out[(blockIdx.x * blockDim.x + threadIdx.x) * H + j] = v;
  }
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}


// MODE 9: the FFMA2 chain of hypothesis j + 1 interleaved one-to-one with the FSET / IADD3 of hypothesis j, the order
// pinned with asm volatile (ptxas otherwise groups the FFMA2s first and leaves runs of ALU instructions at the end).
__device__ __forceinline__ unsigned long long pk(float2 v) { return ((unsigned long long)__float_as_uint(v.y) << 32) | __float_as_uint(v.x); }
__device__ __forceinline__ unsigned long long ffma2v(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned fsetv(unsigned bits, float t) {
  float f;
  asm volatile("set.lt.f32.f32 %0, %1, %2;" : "=f"(f) : "f"(fabsf(__uint_as_float(bits))), "f"(t));
  return __float_as_uint(f);
}
__device__ __forceinline__ unsigned iadd3v(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
template <int H>
__global__ void __launch_bounds__(256, 2) k_score_il(const float4* __restrict__ hyps, int* __restrict__ out, int iters, float t) {
  __shared__ __align__(16) float sx[TP];
  __shared__ __align__(16) float sy[TP];
  __shared__ __align__(16) float sz[TP];
  for (int i = threadIdx.x; i < TP; i += blockDim.x) {
    sx[i] = (i * 37 % 101) * 0.03f; sy[i] = (i * 11 % 97) * 0.031f; sz[i] = (i * 7 % 89) * 0.029f;
  }
  __syncthreads();
  unsigned long long A[H], B[H], C[H], D[H];
  int cnt[H];
  unsigned acc[H];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    float4 h = hyps[(blockIdx.x * blockDim.x + threadIdx.x) * H + j];
    A[j] = pk(make_float2(h.x, h.x)); B[j] = pk(make_float2(h.y, h.y)); C[j] = pk(make_float2(h.z, h.z)); D[j] = pk(make_float2(h.w, h.w));
    cnt[j] = 0; acc[j] = 0;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int pb = 0; pb < TP; pb += 4 * 64) {
#pragma unroll
      for (int j = 0; j < H; ++j) { cnt[j] += ((acc[j] >> 23) * 383u) & 511u; acc[j] = 0; }
      unsigned long long p0 = 0x7FC000007FC00000ull, p1 = p0;  // residual pairs of the previous hypothesis (NaN: never counted)
#pragma unroll 4
      for (int p = pb; p < pb + 4 * 64; p += 4) {
        const float4 X = *reinterpret_cast<const float4*>(&sx[p]);
        const float4 Y = *reinterpret_cast<const float4*>(&sy[p]);
        const float4 Z = *reinterpret_cast<const float4*>(&sz[p]);
        const unsigned long long X0 = pk(make_float2(X.x, X.y)), X1 = pk(make_float2(X.z, X.w));
        const unsigned long long Y0 = pk(make_float2(Y.x, Y.y)), Y1 = pk(make_float2(Y.z, Y.w));
        const unsigned long long Z0 = pk(make_float2(Z.x, Z.y)), Z1 = pk(make_float2(Z.z, Z.w));
#pragma unroll
        for (int j = 0; j < H; ++j) {
          const int q = (j + H - 1) % H;  // the hypothesis whose residuals are compared now
          unsigned f0 = 0, f1 = 0, f2 = 0, f3 = 0;
          unsigned long long r0 = ffma2v(C[j], Z0, D[j]);
          f0 = fsetv((unsigned)p0, t);
          unsigned long long r1 = ffma2v(C[j], Z1, D[j]);
          f1 = fsetv((unsigned)(p0 >> 32), t);
          r0 = ffma2v(B[j], Y0, r0);
          f2 = fsetv((unsigned)p1, t);
          r1 = ffma2v(B[j], Y1, r1);
          f3 = fsetv((unsigned)(p1 >> 32), t);
          r0 = ffma2v(A[j], X0, r0);
          acc[q] = iadd3v(acc[q], f0, f1);
          r1 = ffma2v(A[j], X1, r1);
          acc[q] = iadd3v(acc[q], f2, f3);
          p0 = r0; p1 = r1;
        }
      }
      {  // drain: the last hypothesis of the last step
        const unsigned f0 = fsetv((unsigned)p0, t), f1 = fsetv((unsigned)(p0 >> 32), t), f2 = fsetv((unsigned)p1, t), f3 = fsetv((unsigned)(p1 >> 32), t);
        acc[H - 1] = iadd3v(acc[H - 1], f0, f1);
        acc[H - 1] = iadd3v(acc[H - 1], f2, f3);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < H; ++j) out[(blockIdx.x * blockDim.x + threadIdx.x) * H + j] = cnt[j] + (int)(((acc[j] >> 23) * 383u) & 511u);
}
template <int H>
void run_score_il(const char* name, int nsm, float4* dh, int* dout, double peak_tf) {
  int grid = nsm * 2, iters = 64;
  float ms = time_ms([&] { k_score_il<H><<<grid, 256>>>(dh, dout, iters, 0.1f); });
  CK(cudaGetLastError());
  double pairs = (double)grid * 256 * H * (double)TP * iters;
  double tf = pairs * 6 / (ms * 1e-3) / 1e12;
  printf("%-34s H=%d  %8.3f ms  %8.3f Tpairs/s  %7.2f TFLOP/s-equiv  %5.1f%% of ffma peak\n", name, H, ms, pairs / (ms * 1e-3) / 1e12, tf, 100 * tf / peak_tf);
}

template <int H, int MODE>
void run_score(const char* name, int nsm, float4* dh, int* dout, double peak_tf) {
  int grid = nsm * 2, iters = 64;
  float ms = time_ms([&] { k_score<H, MODE><<<grid, 256>>>(dh, dout, iters, 0.1f); });
  CK(cudaGetLastError());
  double pairs = (double)grid * 256 * H * (double)TP * iters;
  double tf = pairs * 6 / (ms * 1e-3) / 1e12;
  printf("%-34s H=%d  %8.3f ms  %8.3f Tpairs/s  %7.2f TFLOP/s-equiv  %5.1f%% of ffma peak\n", name, H, ms, pairs / (ms * 1e-3) / 1e12, tf, 100 * tf / peak_tf);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, nsm);
  float* dout; CK(cudaMalloc(&dout, sizeof(float) * nsm * 8 * 1024));
  int iters = 4096;
  double best_tf = 0;
  for (int cps = 2; cps <= 8; cps *= 2) {
    int grid = nsm * cps;
    float ms = time_ms([&] { k_ffma<<<grid, 256>>>(dout, iters, 1.0001f, 0.5f); });
    double fl = (double)grid * 256 * iters * 8 * 16 * 2;
    double tf = fl / (ms * 1e-3) / 1e12; if (tf > best_tf) best_tf = tf;
    printf("FFMA  peak grid=%d x256: %.3f ms  %.2f TFLOP/s\n", grid, ms, tf);
    ms = time_ms([&] { k_ffma2<<<grid, 256>>>(dout, iters, 1.0001f, 0.5f); });
    fl = (double)grid * 256 * iters * 8 * 8 * 4;
    tf = fl / (ms * 1e-3) / 1e12;
    printf("FFMA2 peak grid=%d x256: %.3f ms  %.2f TFLOP/s\n", grid, ms, tf);
  }
  {
    int grid = nsm * 4;
    float ms = time_ms([&] { k_fset<<<grid, 256>>>(dout, iters, 0.5f); });
    double ops = (double)grid * 256 * iters * 8 * 16;
    printf("FSET.BF rate: %.3f ms  %.2f Tops/s  (%.1f lanes/clk/SM @1.965GHz)\n", ms, ops / (ms * 1e-3) / 1e12, ops / (ms * 1e-3) / 1.965e9 / nsm);
    ms = time_ms([&] { k_iadd3<<<grid, 256>>>((int*)dout, iters, 3, 5); });
    printf("IADD3 rate: %.3f ms  %.2f Tops/s  (%.1f lanes/clk/SM @1.965GHz)\n", ms, ops / (ms * 1e-3) / 1e12, ops / (ms * 1e-3) / 1.965e9 / nsm);
  }
  run_mix<0, true>(nsm, dout); run_mix<4, true>(nsm, dout); run_mix<8, true>(nsm, dout); run_mix<12, true>(nsm, dout);
  run_mix<0, false>(nsm, dout); run_mix<4, false>(nsm, dout); run_mix<8, false>(nsm, dout); run_mix<12, false>(nsm, dout);
  size_t nh = (size_t)nsm * 2 * 256 * 8;
  std::vector<float4> hh(nh);
  for (size_t i = 0; i < nh; ++i) { hh[i] = make_float4(0.6f, 0.48f, 0.64f, -(float)(i % 100) * 0.03f); }
  float4* dh; CK(cudaMalloc(&dh, nh * sizeof(float4))); CK(cudaMemcpy(dh, hh.data(), nh * sizeof(float4), cudaMemcpyHostToDevice));
  int* dcnt; CK(cudaMalloc(&dcnt, nh * sizeof(int)));
  run_score<8, 0>("scalar FFMA + int count", nsm, dh, dcnt, best_tf);
  run_score<4, 0>("scalar FFMA + int count", nsm, dh, dcnt, best_tf);
  run_score<8, 1>("FFMA2 + int count", nsm, dh, dcnt, best_tf);
  run_score<4, 1>("FFMA2 + int count", nsm, dh, dcnt, best_tf);
  run_score<8, 2>("FFMA2 + float count (FADD2)", nsm, dh, dcnt, best_tf);
  run_score<4, 2>("FFMA2 + float count (FADD2)", nsm, dh, dcnt, best_tf);
  run_score<8, 3>("FFMA2 + mask popc", nsm, dh, dcnt, best_tf);
  run_score<8, 4>("FFMA2 + 3-input add", nsm, dh, dcnt, best_tf);
  run_score<8, 5>("FFMA2 + asm pred add", nsm, dh, dcnt, best_tf);
  run_score<8, 7>("FFMA2 + FSET.BF + IADD3 2-in mod512", nsm, dh, dcnt, best_tf);
  run_score<4, 7>("FFMA2 + FSET.BF + IADD3 2-in mod512", nsm, dh, dcnt, best_tf);
  run_score<8, 8>("FFMA + FSET.BF + IADD3 2-in mod512", nsm, dh, dcnt, best_tf);
  run_score<4, 8>("FFMA + FSET.BF + IADD3 2-in mod512", nsm, dh, dcnt, best_tf);
  run_score_il<8>("FFMA2/FSET/IADD3 interleaved 1:1 (asm volatile)", nsm, dh, dcnt, best_tf);
  {
    std::vector<int> a(16), b(16);
    run_score<8, 7>("  (check) mod512 reference", nsm, dh, dcnt, best_tf);
    CK(cudaMemcpy(a.data(), dcnt, 64, cudaMemcpyDeviceToHost));
    run_score_il<8>("  (check) interleaved", nsm, dh, dcnt, best_tf);
    CK(cudaMemcpy(b.data(), dcnt, 64, cudaMemcpyDeviceToHost));
    printf("interleaved counts %s the reference variant (%d %d %d vs %d %d %d)\n", a == b ? "equal" : "DIFFER FROM", a[0], a[1], a[2], b[0], b[1], b[2]);
  }
  run_score<8, 6>("FFMA2 + FSET mask, IADD3 2-in", nsm, dh, dcnt, best_tf);
  run_score<4, 6>("FFMA2 + FSET mask, IADD3 2-in", nsm, dh, dcnt, best_tf);
  std::vector<int> hc(16); CK(cudaMemcpy(hc.data(), dcnt, 64, cudaMemcpyDeviceToHost));
  printf("check cnt[0..3] = %d %d %d %d\n", hc[0], hc[1], hc[2], hc[3]);
  return 0;
}
