// pr_api.cpp — the extern "C" boundary (include/plane_ransac.h) and the per-call orchestration:
// host draws PCL's sample stream, the device scores it, the host replays PCL's sequential decisions,
// the device refits and peels.  NCCL (loaded at run time) sums counts and moments when the cloud is
// sharded over ranks.  There is no CPU implementation of the device work in this file or anywhere
// in the product: without a CUDA device every entry point that needs one fails.
#include "../../include/plane_ransac.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pr_draw.h"
#include "pr_host.hpp"
#include "pr_kernels.h"

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define PR_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) return fail(PR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define PR_TRY(expr)              \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != PR_OK) return rc__; \
  } while (0)

// ---- NCCL, resolved at run time so the library loads on hosts without it ------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return PR_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(PR_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
  NcclApi a;
  a.handle = h;
#define PR_SYM(field, name)                                                        \
  *(void**)(&a.field) = dlsym(h, name);                                             \
  if (!a.field) return fail(PR_ERR_COMM, "libnccl is missing symbol %s", name)
  PR_SYM(GetUniqueId, "ncclGetUniqueId");
  PR_SYM(CommInitRank, "ncclCommInitRank");
  PR_SYM(CommDestroy, "ncclCommDestroy");
  PR_SYM(AllReduce, "ncclAllReduce");
  PR_SYM(AllGather, "ncclAllGather");
  PR_SYM(GetErrorString, "ncclGetErrorString");
#undef PR_SYM
  g_nccl = a;
  return PR_OK;
}

#define PR_NCCL(expr)                                                                                  \
  do {                                                                                                 \
    ncclResult_t r__ = (expr);                                                                         \
    if (r__ != ncclSuccess) return fail(PR_ERR_COMM, "%s failed: %s", #expr, g_nccl.GetErrorString(r__)); \
  } while (0)

struct HostTimer {
  double* acc;
  std::chrono::steady_clock::time_point t0;
  explicit HostTimer(double* a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~HostTimer() { *acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

enum KClass { KC_STAGE = 0, KC_MODELS, KC_SCORE, KC_REFIT, KC_COMPACT, KC_OTHER, KC_COUNT };

struct TimedSpan {
  cudaEvent_t a, b;
  int cls;
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
};

template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct plane_ransac_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // index lists go back to pinned host buffers while later rounds compute
  cudaEvent_t copy_ready = nullptr;

  // staged cloud (immutable) and the two peel buffers
  pr::CloudView staged, work[2];
  DevBuf<float> staged_mem, work_mem[2];
  DevBuf<int32_t> work_orig[2];
  size_t n_staged = 0;
  bool have_cloud = false;
  pr::CloudView current;
  size_t n_current = 0;
  int scale_exp = 0;
  uint32_t bbox_keys[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0};         // this rank's points
  uint32_t bbox_keys_global[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0};  // all ranks
  // hierarchical scorer: Morton-sorted copies (staged + two peel buffers), block boxes, sort scratch
  DevBuf<float> sorted_mem[3];
  pr::CloudView sorted_view[3];
  bool sorted_staged_valid = false;
  DevBuf<float4> d_bounds;
  DevBuf<float2> d_aux;
  DevBuf<uint32_t> d_keys, d_vals;
  DevBuf<unsigned char> d_sort_temp;
  DevBuf<long long> d_totals2;
  float cmax = 0.f;
  std::vector<size_t> last_offsets;   // plane offsets into d_inl_orig of the last extract call
  std::vector<float> last_coeffs;     // 4 per plane
  DevBuf<int32_t> d_stage_map;  // staged point -> index in the caller's array (only after a filtered staging / restage)
  DevBuf<int32_t> d_stage_map_tmp;
  bool have_stage_map = false;

  DevBuf<float4> aos;        // AoS staging (upload / download)
  DevBuf<uint32_t> d_bbox;   // 6 keys
  // per-call hypothesis state (capacity in draws)
  size_t draw_cap = 0;
  DevBuf<int32_t> d_triples, d_counts, d_good;
  DevBuf<int4> d_sample_pts;
  DevBuf<float4> d_hyps;
  PinBuf<int32_t> h_triples, h_counts, h_good;
  DevBuf<pr::RefitOut> d_refit;
  PinBuf<pr::RefitOut> h_refit;
  DevBuf<long long> d_totals;     // [0..2) local totals, [2 .. 2 + 2*ranks) gathered
  PinBuf<long long> h_totals;
  DevBuf<unsigned char> d_scratch;
  DevBuf<float> d_pcl_sums;  // PR_REFIT_PCL_FLOAT: the nine sequential FP32 sums
  DevBuf<int32_t> d_inl_cur, d_inl_orig;
  PinBuf<float4> h_small;  // raw coefficient read-back
  DevBuf<float4> d_flush;

  // batch of small clouds
  size_t batch_clouds = 0, batch_n = 0, batch_stride = 0;
  DevBuf<float> batch_mem;
  pr::CloudView batch_view;
  std::vector<int> batch_scale_exp;
  DevBuf<uint32_t> d_batch_bbox;
  DevBuf<int32_t> d_batch_idx;       // best triples (3 per cloud) / model index
  DevBuf<double> d_batch_scale;
  DevBuf<pr::RefitOut> d_batch_refit;
  DevBuf<float4> d_batch_hyps;       // best / refined hypothesis per cloud
  DevBuf<int4> d_batch_pts;
  DevBuf<int32_t> d_batch_cnt;
  // device-driven batch path: per-cloud winner, raw model, list offsets, the lists; grid exponents uploaded at staging
  DevBuf<int32_t> d_batch_sexp, d_batch_lists;
  DevBuf<unsigned long long> d_batch_offs;
  DevBuf<uint32_t> d_batch_out;    // flag | best | best count | final count | raw | refined, read back in one copy
  PinBuf<uint32_t> h_batch_out;
  DevBuf<int32_t> d_batch_tri;
  size_t batch_tri_dev_valid = 0;  // number of triple entries of batch_tri that are on the device
  std::vector<int32_t> batch_tri;  // the K triples every cloud of the batch draws (same size, same seed)
  std::vector<cudaEvent_t> head_ev;  // per queued round: its head (draws, sample exchange, models) is done
  std::vector<unsigned long long> timeline;  // last extract call: kStampSlots stamps per device-loop round
  size_t batch_tri_n = 0;
  unsigned batch_tri_seed = 0;

  // postProcessPlanes re-absorption scratch
  DevBuf<pr::ReabsorbPlane> d_rb_planes;
  DevBuf<float4> d_rb_border, d_rb_edges, d_rb_rays;
  DevBuf<int32_t> d_rb_ray_edges, d_rb_counts, d_rb_out_cur, d_rb_out_orig;
  DevBuf<unsigned long long> d_rb_counters, d_rb_keys;
  DevBuf<uint2> d_rb_cand;
  DevBuf<uint32_t> d_rb_claimed;
  DevBuf<unsigned char> d_rb_temp;

  // normal estimation scratch
  DevBuf<unsigned long long> d_nrm_keys;
  DevBuf<uint32_t> d_nrm_idx;
  DevBuf<float> d_nrm_xyz;
  DevBuf<float4> d_nrm_out;
  DevBuf<int32_t> d_nrm_cnt;
  DevBuf<unsigned char> d_nrm_temp;

  // sharding
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;
  // peer-memory exchanges (pr_p2p.cu): this rank's mailbox, the peers' mailboxes mapped through CUDA IPC
  bool p2p_on = false;
  unsigned char* p2p_mailbox = nullptr;
  pr::P2PView p2p_view{};
  DevBuf<unsigned long long> d_p2p_epoch;           // per channel, identical on every rank; advanced by the exchange kernels
  DevBuf<unsigned> d_p2p_aux;                       // [4] timeout flag
  DevBuf<unsigned long long> d_p2p_wait;            // per channel: ns spent waiting for the peers' flags, exchanges
  bool comm_failed = false;                         // an exchange timed out: the ranks are out of step for good
  PinBuf<unsigned> h_p2p_err;
  long long n_global_staged = 0, first_staged = 0, n_global_current = 0, first_current = 0;
  bool global_valid = false;

  pr::IndexSampler sampler;
  int round_loop = PR_LOOP_AUTO;

  // peel loop without the host (run_chain): device-resident round state, per-round records, the sampler's scratch
  DevBuf<pr::RoundState> d_state;
  PinBuf<pr::RoundState> h_state;
  PinBuf<pr::RoundRecord> h_recs;  // page-locked and device-visible: the kernels write the records straight into it
  DevBuf<unsigned> d_chain_tickets;
  std::vector<cudaEvent_t> round_ev;
  DevBuf<uint32_t> d_rnd;  // mt19937(seed) >> 1, the stream every round's draws consume
  uint32_t rnd_seed = 0;
  size_t rnd_count = 0;
  DevBuf<unsigned long long> d_draw_table;
  uint32_t draw_epoch = 0;      // epoch tag of the sampler's table slots; 0: table not initialised
  size_t draw_epoch_slots = 0;
  DevBuf<uint32_t> d_draw_coll;  // kDrawCollCap entries + the counter

  // chunked upload queued by plane_ransac_set_cloud_async: copies run on copy_stream, chunk k is complete at ev[k];
  // the first scoring pass (or any other call that needs the cloud) stages and consumes the chunks as they arrive
  struct PendingUpload {
    bool active = false;
    const pr_point* host = nullptr;
    size_t n = 0;
    int n_chunks = 0, next = 0;
    std::vector<size_t> off;  // chunk k = points [off[k], off[k + 1]), boundaries on whole tiles
    std::vector<cudaEvent_t> ev;
  } pend;
  PinBuf<int4> h_sample_pts;

  // measurement
  cudaEvent_t timer_a = nullptr, timer_b = nullptr;
  bool profiling = false;
  std::vector<TimedSpan> spans;
  std::vector<cudaEvent_t> event_pool;  // recycled by collect_spans: event creation costs microseconds
  pr_profile prof;
};

namespace {

template <typename T>
int dev_reserve(DevBuf<T>& b, size_t n, bool keep = false, cudaStream_t s = nullptr) {
  if (n <= b.cap && b.p) return PR_OK;
  size_t want = std::max(n, (size_t)16);
  T* np = nullptr;
  if (cudaMalloc(&np, want * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return fail(PR_ERR_OOM, "cudaMalloc of %zu bytes failed", want * sizeof(T));
  }
  if (b.p) {
    if (keep && b.cap) {
      if (cudaMemcpyAsync(np, b.p, b.cap * sizeof(T), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
          cudaStreamSynchronize(s) != cudaSuccess)
        return fail(PR_ERR_CUDA, "device buffer grow copy failed");
    }
    cudaFree(b.p);
  }
  b.p = np;
  b.cap = want;
  return PR_OK;
}

template <typename T>
void dev_free(DevBuf<T>& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

template <typename T>
int pin_reserve(PinBuf<T>& b, size_t n, bool keep = false) {
  if (n <= b.cap && b.p) return PR_OK;
  size_t want = std::max(n, (size_t)16);
  T* np = nullptr;
  if (cudaMallocHost(&np, want * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return fail(PR_ERR_OOM, "cudaMallocHost of %zu bytes failed", want * sizeof(T));
  }
  if (b.p) {
    if (keep && b.cap) std::memcpy(np, b.p, b.cap * sizeof(T));
    cudaFreeHost(b.p);
  }
  b.p = np;
  b.cap = want;
  return PR_OK;
}

template <typename T>
void pin_free(PinBuf<T>& b) {
  if (b.p) cudaFreeHost(b.p);
  b.p = nullptr;
  b.cap = 0;
}

cudaEvent_t take_event(plane_ransac_ctx* c) {
  cudaEvent_t e = nullptr;
  if (!c->event_pool.empty()) {
    e = c->event_pool.back();
    c->event_pool.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}

// RAII span: records CUDA events around a kernel class when profiling, and counts launches.
struct Span {
  plane_ransac_ctx* c;
  int cls;
  cudaEvent_t a = nullptr, b = nullptr;
  Span(plane_ransac_ctx* ctx, int k, int launches) : c(ctx), cls(k) {
    long long* lc[KC_COUNT] = {&c->prof.launches_stage, &c->prof.launches_models, &c->prof.launches_score,
                               &c->prof.launches_refit, &c->prof.launches_compact, &c->prof.launches_other};
    *lc[k] += launches;
    if (c->profiling) {
      a = take_event(c);
      b = take_event(c);
      cudaEventRecord(a, c->stream);
    }
  }
  ~Span() {
    if (c->profiling) {
      cudaEventRecord(b, c->stream);
      c->spans.push_back({a, b, cls});
    }
  }
};

void collect_spans(plane_ransac_ctx* c) {
  if (c->spans.empty()) return;
  cudaStreamSynchronize(c->stream);
  double* ms[KC_COUNT] = {&c->prof.ms_stage, &c->prof.ms_models, &c->prof.ms_score,
                          &c->prof.ms_refit, &c->prof.ms_compact, &c->prof.ms_other};
  for (auto& s : c->spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) *ms[s.cls] += t;
    c->event_pool.push_back(s.a);
    c->event_pool.push_back(s.b);
  }
  c->spans.clear();
}

int sync_stream(plane_ransac_ctx* c) {
  HostTimer ht(&c->prof.host_ms_wait);
  PR_CUDA(cudaStreamSynchronize(c->stream));
  return PR_OK;
}

int ensure_staged(plane_ransac_ctx* c);

// keep_pending: the caller deals with a queued asynchronous upload itself (the scoring pass that overlaps it, or a
// new set_cloud that replaces it); every other entry point first lets the upload land.
int check_ctx(plane_ransac_ctx* c, bool keep_pending = false) {
  if (!c) return fail(PR_ERR_INVALID, "null context");
  PR_CUDA(cudaSetDevice(c->device));
  if (!keep_pending && c->pend.active) PR_TRY(ensure_staged(c));
  return PR_OK;
}

int check_params(const pr_params* p) {
  if (!p) return fail(PR_ERR_INVALID, "null params");
  if (!(p->distance_threshold > 0.0) || !std::isfinite(p->distance_threshold))
    return fail(PR_ERR_INVALID, "distance_threshold must be finite and > 0");
  if (p->max_iterations < 0) return fail(PR_ERR_INVALID, "max_iterations must be >= 0");
  if (!(p->probability > 0.0) || p->probability > 1.0) return fail(PR_ERR_INVALID, "probability must be in (0, 1]");
  if (p->dot_order != PR_DOT_PCL_SSE2 && p->dot_order != PR_DOT_FMA) return fail(PR_ERR_INVALID, "unknown dot_order");
  if (p->max_planes < 0) return fail(PR_ERR_INVALID, "max_planes must be >= 0");
  if (p->scorer != PR_SCORER_BRUTE && p->scorer != PR_SCORER_HIER) return fail(PR_ERR_INVALID, "unknown scorer");
  if (p->refit_mode != PR_REFIT_FIXED && p->refit_mode != PR_REFIT_PCL_FLOAT) return fail(PR_ERR_INVALID, "unknown refit_mode");
  return PR_OK;
}

pr::CloudView planes_view(float* base, int32_t* orig, size_t cap) {
  pr::CloudView v;
  v.x = base;
  v.y = base + cap;
  v.z = base + 2 * cap;
  v.orig = orig;
  v.cap = cap;
  return v;
}

int reserve_draws(plane_ransac_ctx* c, size_t need, bool keep) {
  if (need <= c->draw_cap) return PR_OK;
  size_t cap = std::max(need, c->draw_cap * 2);
  PR_TRY(dev_reserve(c->d_triples, 3 * cap, keep, c->stream));
  PR_TRY(dev_reserve(c->d_sample_pts, 3 * cap, keep, c->stream));
  PR_TRY(dev_reserve(c->d_hyps, cap, keep, c->stream));
  PR_TRY(dev_reserve(c->d_counts, cap, keep, c->stream));
  PR_TRY(dev_reserve(c->d_good, cap, keep, c->stream));
  PR_TRY(pin_reserve(c->h_triples, 3 * cap, keep));
  PR_TRY(pin_reserve(c->h_counts, cap, keep));
  PR_TRY(pin_reserve(c->h_good, cap, keep));
  PR_TRY(pin_reserve(c->h_sample_pts, 3 * cap, keep));
  c->draw_cap = cap;
  return PR_OK;
}

int reserve_small(plane_ransac_ctx* c) {
  PR_TRY(dev_reserve(c->d_refit, 1));
  PR_TRY(pin_reserve(c->h_refit, 1));
  PR_TRY(dev_reserve(c->d_totals, 2 + 2 * (size_t)std::max(1, c->n_ranks)));
  PR_TRY(pin_reserve(c->h_totals, 2 + 2 * (size_t)std::max(1, c->n_ranks)));
  PR_TRY(pin_reserve(c->h_small, 4));
  PR_TRY(dev_reserve(c->d_bbox, 8));
  return PR_OK;
}

// Global size / offset / bounding box of a sharded staged cloud.
int refresh_global(plane_ransac_ctx* c) {
  if (!c->have_cloud) return PR_OK;
  if (!c->comm) {
    c->n_global_staged = (long long)c->n_staged;
    c->first_staged = 0;
    std::memcpy(c->bbox_keys_global, c->bbox_keys, sizeof(c->bbox_keys));
    c->scale_exp = pr::scale_exp_from_bbox_keys(c->bbox_keys);
    c->global_valid = true;
    return PR_OK;
  }
  PR_TRY(reserve_small(c));
  // allgather of local sizes
  long long* d = c->d_totals.p;
  c->h_totals.p[0] = (long long)c->n_staged;
  PR_CUDA(cudaMemcpyAsync(d, c->h_totals.p, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  PR_NCCL(g_nccl.AllGather(d, d + 2, 1, ncclInt64, c->comm, c->stream));
  // bounding box: min over keys[0..3), max over keys[3..6)
  PR_CUDA(cudaMemcpyAsync(c->d_bbox.p, c->bbox_keys, 6 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  PR_NCCL(g_nccl.AllReduce(c->d_bbox.p, c->d_bbox.p, 3, ncclUint32, ncclMin, c->comm, c->stream));
  PR_NCCL(g_nccl.AllReduce(c->d_bbox.p + 3, c->d_bbox.p + 3, 3, ncclUint32, ncclMax, c->comm, c->stream));
  uint32_t keys[6];
  PR_CUDA(cudaMemcpyAsync(c->h_totals.p + 2, d + 2, c->n_ranks * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  PR_CUDA(cudaMemcpyAsync(keys, c->d_bbox.p, sizeof(keys), cudaMemcpyDeviceToHost, c->stream));
  PR_CUDA(cudaStreamSynchronize(c->stream));
  long long first = 0, total = 0;
  for (int r = 0; r < c->n_ranks; ++r) {
    if (r == c->rank) first = total;
    total += c->h_totals.p[2 + r];
  }
  c->n_global_staged = total;
  c->first_staged = first;
  std::memcpy(c->bbox_keys_global, keys, sizeof(keys));
  c->scale_exp = pr::scale_exp_from_bbox_keys(keys);
  c->global_valid = true;
  return PR_OK;
}

int stage_from_device(plane_ransac_ctx* c, const float4* d_aos, size_t n, unsigned flags = 0, size_t* n_kept = nullptr,
                      float* centroid_out = nullptr) {
  const size_t cap = pr::padded_capacity(n);
  PR_TRY(dev_reserve(c->staged_mem, 3 * cap));
  PR_TRY(reserve_small(c));
  c->staged = planes_view(c->staged_mem.p, nullptr, cap);
  c->have_stage_map = false;
  c->last_offsets.clear();
  c->last_coeffs.clear();
  c->sorted_staged_valid = false;
  size_t n_out = n;
  if (flags & PR_STAGE_REMOVE_NONFINITE) {
    // stage into a scratch cloud, then compact the finite points into the staged planes (order preserved)
    PR_TRY(dev_reserve(c->work_mem[0], 3 * cap));
    PR_TRY(dev_reserve(c->d_stage_map, cap));
    PR_TRY(dev_reserve(c->d_scratch, pr::compact_scratch_bytes(n) + 64));
    pr::CloudView tmp = planes_view(c->work_mem[0].p, nullptr, cap);
    pr::CloudView dst = c->staged;
    dst.orig = c->d_stage_map.p;
    {
      Span sp(c, KC_STAGE, 2 + (n ? 1 : 0));
      pr::launch_bbox_init(c->d_bbox.p, c->stream);
      pr::launch_stage(d_aos, n, tmp, c->d_bbox.p, c->stream);
      pr::Plane4 none = {0, 0, 0, 0};
      pr::launch_compact(tmp, n, none, 0.f, 2, dst, true, nullptr, nullptr, c->d_scratch.p, c->d_totals.p, c->stream);
    }
    PR_CUDA(cudaGetLastError());
    PR_CUDA(cudaMemcpyAsync(c->h_totals.p, c->d_totals.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    PR_CUDA(cudaMemcpyAsync(c->bbox_keys, c->d_bbox.p, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
    n_out = (size_t)c->h_totals.p[0];
    if (n == 0) {  // launch_compact skipped: pad the empty staged cloud
      Span sp(c, KC_STAGE, 1);
      pr::launch_stage(d_aos, 0, c->staged, c->d_bbox.p, c->stream);
    }
    c->have_stage_map = true;
  } else {
    {
      Span sp(c, KC_STAGE, 2);
      pr::launch_bbox_init(c->d_bbox.p, c->stream);
      pr::launch_stage(d_aos, n, c->staged, c->d_bbox.p, c->stream);
    }
    PR_CUDA(cudaGetLastError());
    PR_CUDA(cudaMemcpyAsync(c->bbox_keys, c->d_bbox.p, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
  }
  c->n_staged = n_out;
  c->have_cloud = true;
  c->current = c->staged;
  c->n_current = n_out;
  c->global_valid = false;
  PR_TRY(refresh_global(c));
  float centroid[3] = {0.f, 0.f, 0.f};
  if (flags & PR_STAGE_TRANSLATE_CENTROID) {
    // exact centroid over all ranks' finite points, about the global bbox corner on the refit grid
    float lo[3];
    double lod[3];
    bool any = true;
    for (int a = 0; a < 3; ++a) {
      if (c->bbox_keys_global[a] > c->bbox_keys_global[3 + a]) any = false;
      lo[a] = any ? pr::key_to_float(c->bbox_keys_global[a]) : 0.f;
      lod[a] = (double)lo[a];
    }
    if (any) {
      PR_CUDA(cudaMemsetAsync(c->d_totals.p, 0, 4 * sizeof(long long), c->stream));
      {
        Span sp(c, KC_STAGE, n_out ? 1 : 0);
        pr::launch_centroid(c->staged, n_out, lod, c->scale_exp, c->d_totals.p, c->num_sms, c->stream);
      }
      if (c->comm) PR_NCCL(g_nccl.AllReduce(c->d_totals.p, c->d_totals.p, 4, ncclInt64, ncclSum, c->comm, c->stream));
      PR_CUDA(cudaMemcpyAsync(c->h_totals.p, c->d_totals.p, 4 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
      PR_CUDA(cudaStreamSynchronize(c->stream));
      if (pr::centroid_from_sums(c->h_totals.p, lo, c->scale_exp, centroid)) {
        {
          Span sp(c, KC_STAGE, n_out ? 1 : 0);
          pr::launch_translate(c->staged, n_out, centroid, c->num_sms, c->stream);
        }
        PR_CUDA(cudaGetLastError());
        // x -> fl(x - c) is monotonic, so the translated box is the box of the translated corners
        for (int a = 0; a < 3; ++a) {
          uint32_t* sets[2] = {c->bbox_keys, c->bbox_keys_global};
          for (uint32_t* k : sets) {
            if (k[a] > k[3 + a]) continue;
            k[a] = pr::float_to_key(pr::key_to_float(k[a]) - centroid[a]);
            k[3 + a] = pr::float_to_key(pr::key_to_float(k[3 + a]) - centroid[a]);
          }
        }
        c->scale_exp = pr::scale_exp_from_bbox_keys(c->bbox_keys_global);
      }
    }
  }
  c->n_global_current = c->n_global_staged;
  c->first_current = c->first_staged;
  if (n_kept) *n_kept = n_out;
  if (centroid_out) std::memcpy(centroid_out, centroid, sizeof(centroid));
  return PR_OK;
}

// ---- asynchronous chunked staging ----------------------------------------------------------------------------------
// Stage chunk k of a queued upload on the main stream once its copy has landed.
int stage_pending_chunk(plane_ransac_ctx* c, int k, pr::CloudView* view, size_t* n_chunk) {
  auto& u = c->pend;
  const size_t off = u.off[k];
  const size_t cnt = u.off[k + 1] - off;
  const bool last = k == u.n_chunks - 1;
  pr::CloudView v = c->staged;
  v.x += off; v.y += off; v.z += off;
  v.cap = last ? c->staged.cap - off : cnt;  // only the last chunk writes the NaN padding
  PR_CUDA(cudaStreamWaitEvent(c->stream, u.ev[k], 0));
  {
    Span sp(c, KC_STAGE, 1);
    pr::launch_stage(c->aos.p + off, cnt, v, c->d_bbox.p, c->stream);
  }
  if (view) *view = v;
  if (n_chunk) *n_chunk = cnt;
  return PR_OK;
}

// After the last chunk was staged and the stream synchronised: bounding box -> refit grid.
int finalize_pending(plane_ransac_ctx* c) {
  c->pend.active = false;
  c->pend.host = nullptr;
  c->global_valid = false;
  PR_TRY(refresh_global(c));
  c->n_global_current = c->n_global_staged;
  c->first_current = c->first_staged;
  return PR_OK;
}

int ensure_staged(plane_ransac_ctx* c) {
  auto& u = c->pend;
  if (!u.active) return PR_OK;
  for (; u.next < u.n_chunks; ++u.next) PR_TRY(stage_pending_chunk(c, u.next, nullptr, nullptr));
  PR_CUDA(cudaGetLastError());
  PR_CUDA(cudaMemcpyAsync(c->bbox_keys, c->d_bbox.p, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  PR_CUDA(cudaStreamSynchronize(c->stream));
  return finalize_pending(c);
}

void cancel_pending(plane_ransac_ctx* c) {
  if (!c->pend.active) return;
  cudaStreamSynchronize(c->copy_stream);
  c->pend.active = false;
  c->pend.host = nullptr;
  c->have_cloud = false;
}

// Morton-sorted copy of the staged cloud for the hierarchical scorer (built once per staging).
int ensure_sorted(plane_ransac_ctx* c, int n_buffers) {
  const size_t cap = c->staged.cap, n = c->n_staged;
  for (int i = 0; i < n_buffers; ++i) {
    PR_TRY(dev_reserve(c->sorted_mem[i], 3 * cap));
    c->sorted_view[i] = planes_view(c->sorted_mem[i].p, nullptr, cap);
  }
  PR_TRY(dev_reserve(c->d_bounds, 2 * ((cap + 31) / 32)));
  PR_TRY(dev_reserve(c->d_totals2, 4));
  if (c->sorted_staged_valid) return PR_OK;
  float lo[3] = {0.f, 0.f, 0.f}, extent = 0.f, cmax = 0.f;
  for (int a = 0; a < 3; ++a) {
    if (c->bbox_keys[a] > c->bbox_keys[3 + a]) continue;
    const float l = pr::key_to_float(c->bbox_keys[a]), h = pr::key_to_float(c->bbox_keys[3 + a]);
    lo[a] = l;
    extent = std::max(extent, h - l);
    cmax = std::max(cmax, std::max(std::fabs(l), std::fabs(h)));
  }
  c->cmax = cmax;
  PR_TRY(dev_reserve(c->d_keys, 2 * std::max<size_t>(n, 1)));
  PR_TRY(dev_reserve(c->d_vals, 2 * std::max<size_t>(n, 1)));
  const size_t tb = pr::sort_temp_bytes(std::max<size_t>(n, 1));
  PR_TRY(dev_reserve(c->d_sort_temp, tb + 256));
  {
    Span sp(c, KC_STAGE, 3);
    pr::launch_morton_sort(c->staged, n, lo, extent, c->d_keys.p, c->d_vals.p, c->d_sort_temp.p, tb, c->sorted_view[0], c->stream);
  }
  PR_CUDA(cudaGetLastError());
  c->sorted_staged_valid = true;
  return PR_OK;
}

// ---- per-round exchanges: peer-memory kernels when the mailboxes are mapped, NCCL otherwise ------------------------
// Mailbox layout (bytes): flags[4 channels][8 ranks] u64, then per channel two buffers (epoch parity):
//   samples  int4[3 * kP2PMaxHyps]                (every entry written by the rank that owns the sampled point)
//   counts   int32[8 ranks][kP2PMaxHyps]          (summed in rank order by the consumer)
//   refit    int64[8 ranks][16], totals int64[8 ranks][2]
// A buffer is rewritten two exchanges of its channel later; a peer can only get there after this rank produced the
// exchange in between, which in stream order follows this rank's consumption of the buffer.
constexpr size_t kP2PMaxHyps = 16384;
enum { P2P_SAMPLES = 0, P2P_COUNTS = 1, P2P_REFIT = 2, P2P_TOTALS = 3 };
constexpr size_t kP2PFlagBytes = 4 * pr::kP2PMaxRanks * sizeof(unsigned long long);
constexpr size_t kP2PSamplesBytes = 3 * kP2PMaxHyps * sizeof(int4);
constexpr size_t kP2PCountsSlot = kP2PMaxHyps * sizeof(int32_t);
constexpr size_t kP2PRefitSlot = 16 * sizeof(long long);
constexpr size_t kP2PTotalsSlot = 2 * sizeof(long long);
constexpr size_t kP2POffSamples = kP2PFlagBytes;
constexpr size_t kP2POffCounts = kP2POffSamples + 2 * kP2PSamplesBytes;
constexpr size_t kP2POffRefit = kP2POffCounts + 2 * pr::kP2PMaxRanks * kP2PCountsSlot;
constexpr size_t kP2POffTotals = kP2POffRefit + 2 * pr::kP2PMaxRanks * kP2PRefitSlot;
constexpr size_t kP2PMailboxBytes = kP2POffTotals + 2 * pr::kP2PMaxRanks * kP2PTotalsSlot;

inline size_t p2p_flag_off(int ch) { return (size_t)ch * pr::kP2PMaxRanks * sizeof(unsigned long long); }

// K1a + its exchange: sample points of `n_samples` indices into dsp on every rank.  st: the shard extent comes from the
// device-resident round state and the exchange is skipped once the peel loop has stopped (run_chain); tail: the small
// step that consumes the result, run by the exchange kernel itself (peer-memory mode only: run_chain requires it).
int exchange_samples(plane_ransac_ctx* c, pr::CloudView src, long long first, size_t n_local, const int32_t* dt, int n_samples,
                     int4* dsp, pr::RoundState* st = nullptr, const pr::P2PTail* tail = nullptr) {
  if (c->p2p_on && (size_t)n_samples <= 3 * kP2PMaxHyps) {
    pr::launch_p2p_samples(c->p2p_view, src, first, n_local, dt, n_samples, kP2POffSamples, kP2PSamplesBytes, p2p_flag_off(P2P_SAMPLES),
                           c->d_p2p_epoch.p + P2P_SAMPLES, dsp, c->d_p2p_aux.p + 4, c->stream, st, c->d_p2p_wait.p + 2 * P2P_SAMPLES, tail,
                           reinterpret_cast<unsigned*>(c->d_p2p_epoch.p + 4));
    return PR_OK;
  }
  pr::launch_gather_samples(src, first, n_local, dt, n_samples, dsp, 1, 0, c->stream, false, st);
  if (c->comm) PR_NCCL(g_nccl.AllReduce(dsp, dsp, (size_t)n_samples * 4, ncclInt32, ncclSum, c->comm, c->stream));
  return PR_OK;
}

// counts[0..n) summed over ranks, in place.
int exchange_counts(plane_ransac_ctx* c, int32_t* dc, size_t n, pr::RoundState* st = nullptr, const pr::P2PTail* tail = nullptr) {
  if (!c->comm) return PR_OK;
  if (c->p2p_on && n <= kP2PMaxHyps) {
    pr::launch_p2p_allreduce_i32(c->p2p_view, dc, n, kP2POffCounts, pr::kP2PMaxRanks * kP2PCountsSlot, kP2PCountsSlot, p2p_flag_off(P2P_COUNTS),
                                 c->d_p2p_epoch.p + P2P_COUNTS, dc, c->d_p2p_aux.p + 4, c->stream, st, c->d_p2p_wait.p + 2 * P2P_COUNTS, tail,
                                 reinterpret_cast<unsigned*>(c->d_p2p_epoch.p + 4) + 2);
    return PR_OK;
  }
  PR_NCCL(g_nccl.AllReduce(dc, dc, n, ncclInt32, ncclSum, c->comm, c->stream));
  return PR_OK;
}

// the 16 integer moments of d_refit summed over ranks, in place (the pivot behind them is identical everywhere).
int exchange_refit(plane_ransac_ctx* c, pr::RoundState* st = nullptr, const pr::P2PTail* tail = nullptr) {
  if (!c->comm) return PR_OK;
  if (c->p2p_on) {
    long long* m = reinterpret_cast<long long*>(c->d_refit.p);
    pr::launch_p2p_allreduce_i64(c->p2p_view, m, 16, kP2POffRefit, pr::kP2PMaxRanks * kP2PRefitSlot, kP2PRefitSlot, p2p_flag_off(P2P_REFIT),
                                 c->d_p2p_epoch.p + P2P_REFIT, m, c->d_p2p_aux.p + 4, c->stream, st, c->d_p2p_wait.p + 2 * P2P_REFIT, tail);
    return PR_OK;
  }
  PR_NCCL(g_nccl.AllReduce(c->d_refit.p, c->d_refit.p, 16, ncclInt64, ncclSum, c->comm, c->stream));
  return PR_OK;
}

// all-gather of (remaining, inliers): d_totals[0..2) of every rank -> d_totals[2 + 2r ..).
int exchange_totals(plane_ransac_ctx* c, pr::RoundState* st = nullptr, const pr::P2PTail* tail = nullptr) {
  if (!c->comm) return PR_OK;
  if (c->p2p_on) {
    pr::launch_p2p_allgather_i64(c->p2p_view, c->d_totals.p, 2, kP2POffTotals, pr::kP2PMaxRanks * kP2PTotalsSlot, kP2PTotalsSlot,
                                 p2p_flag_off(P2P_TOTALS), c->d_p2p_epoch.p + P2P_TOTALS, c->d_totals.p + 2, c->d_p2p_aux.p + 4, c->stream, st,
                                 c->d_p2p_wait.p + 2 * P2P_TOTALS, tail);
    return PR_OK;
  }
  PR_NCCL(g_nccl.AllGather(c->d_totals.p, c->d_totals.p + 2, 2, ncclInt64, c->comm, c->stream));
  return PR_OK;
}

// collective: every rank of the communicator is in this call (destroy of a context whose mailboxes are mapped, or the
// agreed failure of p2p_setup)
void p2p_teardown(plane_ransac_ctx* c, bool collective) {
  for (int r = 0; r < pr::kP2PMaxRanks; ++r) {
    if (c->p2p_view.peers[r] && r != c->rank) cudaIpcCloseMemHandle(c->p2p_view.peers[r]);
    c->p2p_view.peers[r] = nullptr;
  }
  // no rank frees its mailbox while a peer may still have it mapped: every rank closes its mappings first, then all
  // meet in one NCCL all-reduce (skipped when an exchange had timed out: the peers cannot be relied on any more)
  if (collective && c->comm && c->n_ranks > 1 && !c->comm_failed && c->d_totals.p && g_nccl.AllReduce) {
    if (g_nccl.AllReduce(c->d_totals.p, c->d_totals.p, 1, ncclInt64, ncclSum, c->comm, c->stream) == ncclSuccess)
      cudaStreamSynchronize(c->stream);
  }
  if (c->p2p_mailbox) cudaFree(c->p2p_mailbox);
  c->p2p_mailbox = nullptr;
  c->p2p_on = false;
  cudaGetLastError();
}

// After a stream synchronisation: did a consumer give up waiting for a peer?
int p2p_check(plane_ransac_ctx* c) {
  if (c->p2p_on && c->h_p2p_err.p && *c->h_p2p_err.p) {
    // the ranks are out of step from here on (the late peer's epochs no longer match): every later collective call
    // on this context fails at once instead of waiting 20 s again or consuming unreduced data
    c->comm_failed = true;
    *c->h_p2p_err.p = 0;
    cudaMemsetAsync(c->d_p2p_aux.p + 4, 0, sizeof(unsigned), c->stream);
    return fail(PR_ERR_COMM, "a peer rank did not deliver its part of an exchange within ~20 s; the communicator is unusable");
  }
  return PR_OK;
}

int check_comm(plane_ransac_ctx* c) {
  if (c->comm && c->comm_failed) return fail(PR_ERR_COMM, "an earlier exchange timed out; destroy this context and create the communicator again");
  return PR_OK;
}

// Bytes one peel launch moves (what the HBM roofline of K5 is computed from): the coordinate planes of the source, its
// original-index plane when it has one (the staged cloud has none: the index is the position), the remaining cloud
// with its index plane, and 4 B per inlier for each list written.
long long compact_bytes(const pr::CloudView& src, size_t n, bool write_remaining, long long n_rem, long long n_inl, bool cur, bool orig) {
  return (src.orig ? 16ll : 12ll) * (long long)n + (write_remaining ? 16ll * n_rem : 0) + ((cur ? 4ll : 0) + (orig ? 4ll : 0)) * n_inl;
}

struct SegmentOut {
  float coeff[4] = {0, 0, 0, 0};
  long long n_inl_local = 0, n_rem_local = 0;
  long long n_inl_global = 0;
  std::vector<long long> rem_per_rank;  // sharded: remaining points per rank after this round
};

// One pcl::SACSegmentation::segment() on the cloud `src` (n_local points here, n_global overall,
// this rank's first global index `first`).  Inlier positions / original indices go to d_inl_*; with
// write_remaining the non-inliers are compacted into dst.
int segment_core(plane_ransac_ctx* c, const pr_params* prm, pr::CloudView src, size_t n_local, long long n_global,
                 long long first, bool write_remaining, pr::CloudView dst, int32_t* d_inl_cur, int32_t* d_inl_orig,
                 pr_segment_info* info, SegmentOut* out, const pr::CloudView* ssrc = nullptr,
                 const pr::CloudView* sdst = nullptr) {
  // ssrc: Morton-sorted copy of src (hierarchical scorer); sdst receives its peeled remainder
  HostTimer whole(&c->prof.host_ms_total);
  PR_TRY(check_comm(c));
  pr_segment_info inf;
  std::memset(&inf, 0, sizeof(inf));
  inf.n_cloud = n_global;
  inf.scale_exp = c->scale_exp;
  *out = SegmentOut();
  out->n_rem_local = (long long)n_local;
  const float t = pr::threshold_up(prm->distance_threshold);
  PR_TRY(reserve_small(c));

  const bool hier = ssrc != nullptr && n_local > 0;
  if (hier) {
    Span sp(c, KC_SCORE, 1);
    pr::launch_block_bounds(*ssrc, n_local, c->d_bounds.p, c->stream);
  }

  pr::RansacReplay replay(std::max(1ll, n_global), prm->max_iterations, prm->probability);
  int total_draws = 0;
  // A cloud still arriving from plane_ransac_set_cloud_async: the first batch of hypotheses is scored chunk by chunk
  // as the copies land (the sample points come from the caller's host buffer), so the upload hides under the scoring.
  const bool overlap = c->pend.active && src.x == c->staged.x && !hier && n_global >= 3 && !replay.done();
  if (c->pend.active && !overlap) PR_TRY(ensure_staged(c));
  if (n_global >= 3 && !replay.done()) {
    pr::IndexSampler& sampler = c->sampler;  // kept across rounds: a reset clears only the touched table slots
    sampler.reset((size_t)n_global, prm->seed);
    sampler.reserve((size_t)std::min<long long>((long long)prm->max_iterations + 1, 1 << 20));
    int prev_batch = 0;
    while (!replay.done()) {
      const long long trials_left = (long long)prm->max_iterations + 1 - replay.iterations();
      long long B;
      if (prm->probability >= 1.0) {
        B = std::min<long long>(trials_left, 1 << 20);  // score-all mode: at most 1M hypotheses per device batch
      } else if (prev_batch == 0) {
        B = std::min<long long>(trials_left, 256);
      } else {
        B = std::min<long long>(trials_left, std::max<long long>(std::min(replay.draws_wanted(), 8192), 2 * prev_batch));
      }
      if (B < 1) B = 1;
      if ((long long)total_draws + B > (long long)INT_MAX / 4) return fail(PR_ERR_INVALID, "too many draws");
      PR_TRY(reserve_draws(c, (size_t)total_draws + (size_t)B, total_draws > 0));
      if (overlap && c->pend.active) {
        auto& u = c->pend;
        int32_t* ht = c->h_triples.p;
        {
          HostTimer tm(&c->prof.host_ms_sampling);
          for (long long j = 0; j < B; ++j) sampler.draw(ht + 3 * j);
        }
        int4* hsp = c->h_sample_pts.p;
        for (long long i = 0; i < 3 * B; ++i) {  // random reads of a buffer far larger than the caches: fetch them all first
          const long long local = (long long)ht[i] - first;
          if (local >= 0 && local < (long long)n_local) __builtin_prefetch(&u.host[local], 0, 0);
        }
        for (long long i = 0; i < 3 * B; ++i) {
          const long long local = (long long)ht[i] - first;  // sharded: only the owner of a point contributes its bits
          int4 v = {0, 0, 0, 0};
          if (local >= 0 && local < (long long)n_local) {
            const pr_point& q = u.host[local];
            std::memcpy(&v.x, &q.x, 4);
            std::memcpy(&v.y, &q.y, 4);
            std::memcpy(&v.z, &q.z, 4);
            v.w = 0x3F800000;
          }
          hsp[i] = v;
        }
        // no cudaMemcpy here: a host-to-device copy would queue behind the cloud's chunks on the copy engine; a
        // kernel reads the (pinned, device-visible) sample points over PCIe instead
        {
          Span sp(c, KC_MODELS, 2);
          pr::launch_copy(reinterpret_cast<const float4*>(hsp), reinterpret_cast<float4*>(c->d_sample_pts.p), (size_t)(3 * B), c->num_sms, c->stream);
          if (c->comm)  // once per extraction, not per round: NCCL is fine here
            PR_NCCL(g_nccl.AllReduce(c->d_sample_pts.p, c->d_sample_pts.p, (size_t)(3 * B) * 4, ncclInt32, ncclSum, c->comm, c->stream));
          pr::launch_models(c->d_sample_pts.p, (int)B, c->d_hyps.p, c->d_good.p, c->stream);
        }
        PR_CUDA(cudaMemsetAsync(c->d_counts.p, 0, B * sizeof(int32_t), c->stream));
        for (; u.next < u.n_chunks;) {
          pr::CloudView view;
          size_t cnt = 0;
          PR_TRY(stage_pending_chunk(c, u.next, &view, &cnt));
          ++u.next;
          Span sp(c, KC_SCORE, 0);
          c->prof.launches_score += pr::launch_score(view, cnt, 1, 0, c->d_hyps.p, (int)B, t, prm->dot_order, c->d_counts.p, c->num_sms, c->stream);
          c->prof.pairs_scored += (long long)cnt * B;
        }
        PR_CUDA(cudaMemcpyAsync(c->bbox_keys, c->d_bbox.p, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      } else {
      // The batch is issued in sub-batches so that the host draws the next triples (a sequential
      // permutation walk, ~0.1 us per draw) while the device is already scoring the previous ones.
      // Two sub-batches: the first is just large enough that scoring it (~N / 6.5e12 s per hypothesis, ~3x less
      // with the hierarchical scorer)
      // takes the device as long as drawing the rest takes the host (~0.08 us per draw); fewer, larger
      // launches score more efficiently and, when sharded, need fewer collectives.
      // Sized from quantities every rank agrees on (the global size and the requested scorer, never the local shard
      // size): each sub-batch issues collectives, so all ranks must cut the batch identically however unevenly the
      // peeled planes leave the shards.
      const double pts_per_rank = (double)std::max<long long>(n_global / std::max(1, c->n_ranks), 1);
      const double dev_s_per_hyp = pts_per_rank / (ssrc != nullptr ? 2.0e13 : 6.5e12), host_s_per_draw = 8e-8;
      long long first_sb = 256;
      while (first_sb < B && (double)first_sb < (double)B * host_s_per_draw / (dev_s_per_hyp + host_s_per_draw)) first_sb *= 2;
      while (first_sb * 4 < B) first_sb *= 2;  // launches below a quarter of the batch score a few per cent less efficiently
      long long prev_sb = 0;
      for (long long done = 0; done < B;) {
        long long sb = prev_sb == 0 ? std::min<long long>(B, first_sb) : B - done;
        if (B - done - sb <= 256) sb = B - done;  // fold a short tail into this sub-batch
        prev_sb = sb;
        const size_t at = (size_t)total_draws + (size_t)done;
        int32_t* ht = c->h_triples.p + 3 * at;
        {
          HostTimer tm(&c->prof.host_ms_sampling);
          for (long long j = 0; j < sb; ++j) sampler.draw(ht + 3 * j);
        }
        int32_t* dt = c->d_triples.p + 3 * at;
        int4* dsp = c->d_sample_pts.p + 3 * at;
        float4* dh = c->d_hyps.p + at;
        int32_t* dc = c->d_counts.p + at;
        int32_t* dg = c->d_good.p + at;
        PR_CUDA(cudaMemcpyAsync(dt, ht, 3 * sb * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        {
          Span sp(c, KC_MODELS, 2);
          PR_TRY(exchange_samples(c, src, first, n_local, dt, (int)(3 * sb), dsp));
          pr::launch_models(dsp, (int)sb, dh, dg, c->stream);
        }
        PR_CUDA(cudaMemsetAsync(dc, 0, sb * sizeof(int32_t), c->stream));
        {
          Span sp(c, KC_SCORE, 0);
          if (hier) {
            PR_TRY(dev_reserve(c->d_aux, c->draw_cap));
            c->prof.launches_score += pr::launch_score_hier(*ssrc, n_local, c->d_bounds.p, dh, c->d_aux.p + at, (int)sb, t, c->cmax,
                                                            prm->dot_order, dc, c->num_sms, c->stream);
          } else {
            c->prof.launches_score += pr::launch_score(src, n_local, 1, 0, dh, (int)sb, t, prm->dot_order, dc, c->num_sms, c->stream);
          }
          c->prof.pairs_scored += (long long)n_local * sb;
        }
        done += sb;
      }
      }
      int32_t* dc = c->d_counts.p + total_draws;
      int32_t* dg = c->d_good.p + total_draws;
      PR_TRY(exchange_counts(c, dc, (size_t)B));
      PR_CUDA(cudaGetLastError());
      PR_CUDA(cudaMemcpyAsync(c->h_counts.p + total_draws, dc, B * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
      PR_CUDA(cudaMemcpyAsync(c->h_good.p + total_draws, dg, B * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
      if (c->p2p_on) PR_CUDA(cudaMemcpyAsync(c->h_p2p_err.p, c->d_p2p_aux.p + 4, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
      PR_TRY(sync_stream(c));
      PR_TRY(p2p_check(c));  // a timed-out exchange left unreduced counts: stop before anything consumes them
      if (c->pend.active) {  // the whole cloud is staged now: bounding box -> refit grid
        PR_TRY(finalize_pending(c));
        inf.scale_exp = c->scale_exp;
      }
      {
        HostTimer tm(&c->prof.host_ms_replay);
        std::vector<uint8_t> good8((size_t)B);
        for (long long j = 0; j < B; ++j) good8[j] = c->h_good.p[total_draws + j] ? 1 : 0;
        replay.feed(c->h_counts.p + total_draws, good8.data(), (int)B);
      }
      total_draws += (int)B;
      prev_batch = (int)B;
    }
  }
  inf.n_scored = total_draws;
  inf.iterations = replay.iterations();
  inf.draws = replay.draws_used();
  inf.skipped = replay.skipped();
  const int best = replay.best_draw();
  if (best < 0) {
    if (info) *info = inf;
    return PR_OK;  // "No solution found": outputs cleared
  }
  inf.ok = 1;
  inf.best_count = replay.best_count();
  inf.n_inliers_raw = inf.best_count;
  for (int i = 0; i < 3; ++i) inf.best_sample[i] = c->h_triples.p[3 * (size_t)best + i];

  float refined[4];
  const bool pcl_float = prm->optimize_coefficients && prm->refit_mode == PR_REFIT_PCL_FLOAT;
  if (pcl_float) {
    // PCL 1.8's own arithmetic: selectWithinDistance(raw model) in index order, nine sequential FP32 sums over that
    // list (one device thread), FP32 eigen33 on the host.  The list is built where the final selection will be written.
    if (c->comm) return fail(PR_ERR_INVALID, "PR_REFIT_PCL_FLOAT needs the whole cloud on one GPU (a sequential sum has no sharded form)");
    if (!d_inl_cur) return fail(PR_ERR_INVALID, "PR_REFIT_PCL_FLOAT needs the inlier list buffer");
    PR_CUDA(cudaMemcpyAsync(c->h_small.p, c->d_hyps.p + best, sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    PR_TRY(sync_stream(c));
    std::memcpy(inf.raw_coeff, c->h_small.p, 4 * sizeof(float));
    PR_TRY(dev_reserve(c->d_scratch, pr::compact_scratch_bytes(n_local) + 64));
    PR_TRY(dev_reserve(c->d_pcl_sums, 16));
    {
      Span sp(c, KC_REFIT, n_local ? 2 : 1);
      const pr::Plane4 raw = {inf.raw_coeff[0], inf.raw_coeff[1], inf.raw_coeff[2], inf.raw_coeff[3]};
      pr::CloudView none;
      pr::launch_compact(src, n_local, raw, t, prm->dot_order, none, false, d_inl_cur, nullptr, c->d_scratch.p, c->d_totals.p, c->stream);
      pr::launch_refit_pcl_float(src, d_inl_cur, c->d_totals.p + 1, c->d_pcl_sums.p, c->stream);
      c->prof.points_refit += (long long)n_local;
      c->prof.bytes_refit += 12ll * (long long)n_local;
    }
    PR_CUDA(cudaGetLastError());
    float sums[9];
    PR_CUDA(cudaMemcpyAsync(sums, c->d_pcl_sums.p, sizeof(sums), cudaMemcpyDeviceToHost, c->stream));
    PR_CUDA(cudaMemcpyAsync(c->h_totals.p, c->d_totals.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    PR_TRY(sync_stream(c));
    std::memcpy(refined, inf.raw_coeff, sizeof(refined));
    pr::plane_from_pcl_float_sums(sums, c->h_totals.p[1], refined);  // < 4 inliers: keeps the raw model
  } else if (prm->optimize_coefficients) {
    PR_CUDA(cudaMemsetAsync(c->d_refit.p, 0, sizeof(pr::RefitOut), c->stream));
    {
      Span sp(c, KC_REFIT, 1);
      pr::launch_refit(src, n_local, c->d_hyps.p, c->d_sample_pts.p, best, t, prm->dot_order, c->scale_exp,
                       c->d_refit.p, c->num_sms, c->stream);
      c->prof.points_refit += (long long)n_local;
      c->prof.bytes_refit += 12ll * (long long)n_local;
    }
    PR_TRY(exchange_refit(c));
    PR_CUDA(cudaMemcpyAsync(c->h_refit.p, c->d_refit.p, sizeof(pr::RefitOut), cudaMemcpyDeviceToHost, c->stream));
  }
  if (!pcl_float) {
    PR_CUDA(cudaMemcpyAsync(c->h_small.p, c->d_hyps.p + best, sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    PR_TRY(sync_stream(c));
    std::memcpy(inf.raw_coeff, c->h_small.p, 4 * sizeof(float));
    std::memcpy(refined, inf.raw_coeff, sizeof(refined));
  }
  if (prm->optimize_coefficients && !pcl_float) {
    int64_t m[16];
    for (int i = 0; i < 16; ++i) m[i] = (int64_t)c->h_refit.p->m[i];
    // sharded: every rank wrote the same pivot; the all-reduce summed only m[]
    float pivot[3] = {c->h_refit.p->pivot[0], c->h_refit.p->pivot[1], c->h_refit.p->pivot[2]};
    pr::plane_from_moments(m, pivot, c->scale_exp, refined);  // < 4 inliers: keeps the raw model
  }

  // final selection with the (refined) coefficients; peel
  PR_TRY(dev_reserve(c->d_scratch, pr::compact_scratch_bytes(n_local) + 64));
  {
    Span sp(c, KC_COMPACT, n_local ? 1 : 0);
    pr::Plane4 pl = {refined[0], refined[1], refined[2], refined[3]};
    pr::launch_compact(src, n_local, pl, t, prm->dot_order, dst, write_remaining, d_inl_cur, d_inl_orig,
                       c->d_scratch.p, c->d_totals.p, c->stream);
    c->prof.points_compact += (long long)n_local;
  }
  PR_CUDA(cudaGetLastError());
  PR_TRY(exchange_totals(c));
  PR_CUDA(cudaMemcpyAsync(c->h_totals.p, c->d_totals.p, (2 + (c->comm ? 2 * c->n_ranks : 0)) * sizeof(long long),
                          cudaMemcpyDeviceToHost, c->stream));
  if (c->p2p_on) PR_CUDA(cudaMemcpyAsync(c->h_p2p_err.p, c->d_p2p_aux.p + 4, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  if (hier && sdst && write_remaining) {
    // the sorted copy is peeled with the same predicate (stable, so it stays in Morton order)
    Span sp(c, KC_COMPACT, 1);
    pr::Plane4 pl = {refined[0], refined[1], refined[2], refined[3]};
    pr::launch_compact(*ssrc, n_local, pl, t, prm->dot_order, *sdst, true, nullptr, nullptr, c->d_scratch.p, c->d_totals2.p, c->stream);
    c->prof.points_compact += (long long)n_local;
    c->prof.bytes_compact += 12ll * (long long)n_local;
  }
  PR_TRY(sync_stream(c));
  PR_TRY(p2p_check(c));
  out->n_rem_local = c->h_totals.p[0];
  out->n_inl_local = c->h_totals.p[1];
  out->n_inl_global = out->n_inl_local;
  if (c->comm) {
    out->n_inl_global = 0;
    out->rem_per_rank.resize(c->n_ranks);
    for (int r = 0; r < c->n_ranks; ++r) {
      out->rem_per_rank[r] = c->h_totals.p[2 + 2 * r];
      out->n_inl_global += c->h_totals.p[2 + 2 * r + 1];
    }
  }
  c->prof.bytes_compact += compact_bytes(src, n_local, write_remaining, out->n_rem_local, out->n_inl_local, d_inl_cur != nullptr, d_inl_orig != nullptr);
  c->prof.points_kept += write_remaining ? out->n_rem_local : 0;
  c->prof.points_peeled += out->n_inl_local;
  std::memcpy(out->coeff, refined, sizeof(refined));
  inf.n_inliers = (int)out->n_inl_global;
  if (info) *info = inf;
  return PR_OK;
}

// Where the peel loop stands: shared by the rounds the device runs on its own (run_chain) and the host-driven ones.
struct PeelCursor {
  int planes = 0;        // planes accepted so far
  pr::CloudView src;     // cloud of the next round
  size_t n_local = 0;
  long long n_global = 0, first = 0;
  size_t off = 0;        // inlier-list entries written so far (this rank)
  int s_cur = 0;         // hierarchical scorer: which sorted copy matches src
};

// The caller's index lists: each plane's segment is copied as soon as its round is complete when the destination is
// pinned.  A destination that turns out too small never interrupts the peel (when sharded every rank has to finish the
// round's collectives): the copies stop and the call reports PR_ERR_CAPACITY once the device state is consistent.
struct ListCopier {
  int32_t* cur = nullptr;
  int32_t* orig = nullptr;
  size_t cap = 0;
  bool overlap = false, overflow = false;
  size_t needed = 0;
  bool wanted() const { return cur || orig; }
  int on_plane(plane_ransac_ctx* c, size_t begin, size_t end, cudaEvent_t ready) {
    if (!wanted()) return PR_OK;
    if (end > cap) {
      overflow = true;
      needed = end;
      return PR_OK;
    }
    if (!overlap || overflow || end == begin) return PR_OK;
    if (ready) PR_CUDA(cudaStreamWaitEvent(c->copy_stream, ready, 0));
    if (cur) PR_CUDA(cudaMemcpyAsync(cur + begin, c->d_inl_cur.p + begin, (end - begin) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy_stream));
    if (orig) PR_CUDA(cudaMemcpyAsync(orig + begin, c->d_inl_orig.p + begin, (end - begin) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy_stream));
    return PR_OK;
  }
};

// ---- the peel loop without the host ----------------------------------------------------------------------------------
// Score-all mode (probability 1: PCL scores max_iterations + 1 samples per round whatever the counts are) with the
// brute-force scorer: nothing the host decides depends on device results, so whole rounds are queued ahead —
// draws, models, scores, computeModel's decision, refit, closed-form plane, peel and the stop rule all run as kernels
// driven by a RoundState in HBM (pr_chain.cu) — and the host only reads each round's record as it completes (to return
// the index lists under the following rounds and to stop queueing once the loop has ended).  Rounds that need PCL's
// redraw rule (a degenerate sample) or whose sampler has too many colliding picks are handed back to the host-driven
// loop (segment_core), which gives the same result.
constexpr int kChainMaxHyps = 16384;  // one exchange kernel per quantity (kP2PMaxHyps), one block-wide decision kernel
constexpr int kChainLookahead = 3;    // rounds queued beyond the last one whose record was read

bool chain_eligible(const plane_ransac_ctx* c, const pr_params* prm, long long n_global) {
  static const bool enabled = [] { const char* e = getenv("PR_CHAIN"); return !(e && atoi(e) == 0); }();
  if (!enabled || c->round_loop == PR_LOOP_HOST || c->pend.active) return false;
  if (!(prm->probability >= 1.0) || prm->scorer != PR_SCORER_BRUTE) return false;
  if (prm->optimize_coefficients && prm->refit_mode != PR_REFIT_FIXED) return false;  // the sequential FP32 sums are host-driven
  if (prm->max_iterations < 1 || (long long)prm->max_iterations + 1 > kChainMaxHyps) return false;
  if (c->comm && !c->p2p_on) return false;
  // about (3K)^2 / N of a round's picks collide; beyond what the device replays the round would come straight back
  const double k = (double)prm->max_iterations + 1.0;
  if (n_global < 3 || 9.0 * k * k / (double)n_global > 0.75 * pr::kDrawMaxCollisions) return false;
  return true;
}

int run_chain(plane_ransac_ctx* c, const pr_params* prm, PeelCursor& cur, float* coeffs, size_t* plane_offsets, pr_segment_info* infos,
              ListCopier& lists, bool* stopped, bool* handed_back) {
  HostTimer whole(&c->prof.host_ms_total);
  const int K = prm->max_iterations + 1;
  const float t = pr::threshold_up(prm->distance_threshold);
  const int max_rounds = prm->max_planes - cur.planes;
  const int planes_at_start = cur.planes;
  const bool sharded = c->comm != nullptr;
  PR_TRY(reserve_draws(c, (size_t)K, false));
  PR_TRY(reserve_small(c));
  PR_TRY(dev_reserve(c->d_state, 1));
  PR_TRY(pin_reserve(c->h_state, 1));
  PR_TRY(pin_reserve(c->h_recs, (size_t)max_rounds));
  PR_TRY(dev_reserve(c->d_chain_tickets, 4));
  while ((int)c->round_ev.size() < max_rounds) {
    cudaEvent_t e = nullptr;
    PR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->round_ev.push_back(e);
  }
  // Sharded runs that return index lists: a plane's lists must not travel while the next round's head runs its
  // exchanges — a system-scope fence issued while a device-to-host DMA saturates PCIe waits tens of microseconds — so
  // the copy of plane r starts once round r + 1 has reached its scoring launch (head_ev) and is over long before that
  // launch ends.
  const bool stagger_lists = sharded && lists.wanted() && lists.overlap;
  while (stagger_lists && (int)c->head_ev.size() < max_rounds) {
    cudaEvent_t e = nullptr;
    PR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->head_ev.push_back(e);
  }
  const size_t slots = pr::draw_table_slots(K);
  PR_TRY(dev_reserve(c->d_draw_table, slots));
  PR_TRY(dev_reserve(c->d_draw_coll, (size_t)pr::kDrawCollCap + 4));
  const size_t scratch_bytes = (pr::compact_scratch_bytes(cur.n_local) + 7) / 8 * 8;
  PR_TRY(dev_reserve(c->d_scratch, scratch_bytes + 64));
  if (!c->d_rnd.p || c->rnd_seed != prm->seed || c->rnd_count < 3 * (size_t)K) {
    // the stream does not depend on the cloud: generated once per seed, staged through the (pinned) triple buffer
    PR_TRY(dev_reserve(c->d_rnd, 3 * (size_t)K));
    pr::fill_rnd_stream(prm->seed, 3 * (size_t)K, reinterpret_cast<uint32_t*>(c->h_triples.p));
    PR_CUDA(cudaMemcpyAsync(c->d_rnd.p, c->h_triples.p, 3 * (size_t)K * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
    c->rnd_seed = prm->seed;
    c->rnd_count = 3 * (size_t)K;
  }
  pr::RoundState* rs = c->d_state.p;
  {
    pr::RoundState h;
    std::memset(&h, 0, sizeof(h));
    h.n_local = (long long)cur.n_local;
    h.n_global = cur.n_global;
    h.first = cur.first;
    h.inl_off = (long long)cur.off;
    h.round = cur.planes;
    h.best = -1;
    *c->h_state.p = h;
    PR_CUDA(cudaMemcpyAsync(rs, c->h_state.p, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    std::memset(c->h_recs.p, 0, (size_t)max_rounds * sizeof(pr::RoundRecord));  // ran = 0 until a round's kernels say otherwise
  }
  size_t n_bound = cur.n_local;  // upper bound of this rank's cloud for the rounds queued from here on
  int launched = 0, consumed = 0;
  bool stop_seen = false;
  std::vector<pr::CloudView> dst_of((size_t)max_rounds);
  while (consumed < launched || (!stop_seen && launched < max_rounds)) {
    while (!stop_seen && launched < max_rounds && launched < consumed + kChainLookahead) {
      const int r = launched, round_index = planes_at_start + r;
      const pr::CloudView src = r == 0 ? cur.src : c->work[(round_index - 1) & 1];
      const pr::CloudView dst = c->work[round_index & 1];
      dst_of[r] = dst;
      pr::RoundRecord* rec = c->h_recs.p + r;  // mapped host memory (unified addressing)
      uint32_t* coll_count = c->d_draw_coll.p + pr::kDrawCollCap;
      {
        Span sp(c, KC_MODELS, 3);
        if (c->draw_epoch == 0 || c->draw_epoch >= 65535u || c->draw_epoch_slots != slots) {
          // epoch-tagged slots: the sampler's table is zeroed once (and when the 16-bit epoch wraps), never per round
          PR_CUDA(cudaMemsetAsync(c->d_draw_table.p, 0, slots * sizeof(unsigned long long), c->stream));
          PR_CUDA(cudaMemsetAsync(coll_count, 0, sizeof(uint32_t), c->stream));
          c->draw_epoch = 0;
          c->draw_epoch_slots = slots;
        }
        ++c->draw_epoch;
        pr::launch_draw(c->d_rnd.p, K, rs, c->d_triples.p, c->d_draw_table.p, slots, c->draw_epoch, c->d_draw_coll.p, coll_count, rec,
                        c->d_counts.p, c->d_refit.p, c->d_scratch.p, scratch_bytes, c->d_chain_tickets.p, c->stream);
        if (sharded) {  // (chain_eligible: sharded implies peer-memory exchanges, which run the consuming step themselves)
          pr::P2PTail tm;
          tm.kind = pr::P2PTail::kModels;
          tm.hyps_out = c->d_hyps.p;
          tm.good_out = c->d_good.p;
          PR_TRY(exchange_samples(c, src, 0, 0, c->d_triples.p, 3 * K, c->d_sample_pts.p, rs, &tm));
        } else {
          pr::launch_gather_models(src, c->d_triples.p, K, c->d_sample_pts.p, c->d_hyps.p, c->d_good.p, rs, c->stream);
        }
      }
      if (stagger_lists) PR_CUDA(cudaEventRecord(c->head_ev[r], c->stream));
      {
        Span sp(c, KC_SCORE, 0);
        c->prof.launches_score += pr::launch_score(src, n_bound, 1, 0, c->d_hyps.p, K, t, prm->dot_order, c->d_counts.p, c->num_sms, c->stream, rs);
      }
      if (sharded) {
        Span sp(c, KC_OTHER, 1);
        pr::P2PTail tr;
        tr.kind = pr::P2PTail::kReplay;
        tr.rec = rec;
        tr.good = c->d_good.p;
        PR_TRY(exchange_counts(c, c->d_counts.p, (size_t)K, rs, &tr));
      } else {
        Span sp(c, KC_OTHER, 1);
        pr::launch_replay(c->d_counts.p, c->d_good.p, K, rs, rec, c->stream);
      }
      // one GPU: the closed-form plane runs in the last block of the refit, the stop rule in the last tile of the peel
      pr::ChainTail tail;
      tail.rec = rec;
      tail.ticket = c->d_chain_tickets.p;
      tail.triples = c->d_triples.p;
      tail.scale_exp = c->scale_exp;
      tail.n_draws = K;
      tail.min_plane = prm->min_plane_size;
      if (prm->optimize_coefficients) {
        {
          Span sp(c, KC_REFIT, 1);
          pr::launch_refit(src, n_bound, c->d_hyps.p, c->d_sample_pts.p, 0, t, prm->dot_order, c->scale_exp, c->d_refit.p, c->num_sms, c->stream, rs,
                           sharded ? nullptr : &tail);
        }
        if (sharded) {
          Span sp(c, KC_OTHER, 1);
          pr::P2PTail tf;
          tf.kind = pr::P2PTail::kFinish;
          tf.rec = rec;
          tf.hyps = c->d_hyps.p;
          tf.triples = c->d_triples.p;
          tf.optimize = 1;
          tf.scale_exp = c->scale_exp;
          tf.n_draws = K;
          PR_TRY(exchange_refit(c, rs, &tf));
        }
      } else {
        Span sp(c, KC_OTHER, 1);
        pr::launch_finish(rs, c->d_hyps.p, c->d_triples.p, c->d_refit.p, 0, c->scale_exp, K, rec, c->stream);
      }
      {
        Span sp(c, KC_COMPACT, 1);
        const pr::Plane4 none = {0, 0, 0, 0};
        pr::launch_compact(src, n_bound, none, t, prm->dot_order, dst, true, c->d_inl_cur.p, c->d_inl_orig.p, c->d_scratch.p, c->d_totals.p,
                           c->stream, nullptr, rs, sharded ? nullptr : &tail);
      }
      if (sharded) {
        Span sp(c, KC_OTHER, 1);
        pr::P2PTail ta;
        ta.kind = pr::P2PTail::kAdvance;
        ta.rec = rec;
        ta.min_plane = prm->min_plane_size;
        PR_TRY(exchange_totals(c, rs, &ta));
      }
      PR_CUDA(cudaGetLastError());
      PR_CUDA(cudaEventRecord(c->round_ev[r], c->stream));
      ++launched;
    }
    {
      HostTimer ht(&c->prof.host_ms_wait);
      PR_CUDA(cudaEventSynchronize(c->round_ev[consumed]));
    }
    const pr::RoundRecord rec = c->h_recs.p[consumed];
    const int r = consumed++;
    if (stop_seen || !rec.ran) {  // queued behind the end of the loop: nothing ran
      stop_seen = true;
      continue;
    }
    if (rec.stop == 2) {  // the round goes back to the host-driven loop; the state is that of its start
      stop_seen = true;
      *handed_back = true;
      continue;
    }
    pr_segment_info inf;
    std::memset(&inf, 0, sizeof(inf));
    inf.n_cloud = rec.n_cloud;
    inf.scale_exp = c->scale_exp;
    inf.ok = rec.ok;
    if (rec.ok) {
      inf.n_scored = inf.iterations = inf.draws = rec.n_draws;
      inf.best_count = inf.n_inliers_raw = rec.best_count;
      for (int i = 0; i < 3; ++i) inf.best_sample[i] = rec.best_sample[i];
      std::memcpy(inf.raw_coeff, rec.raw, sizeof(rec.raw));
      inf.n_inliers = (int)rec.n_inl_global;
      const pr::CloudView src = r == 0 ? cur.src : dst_of[r - 1];
      c->prof.pairs_scored += rec.n_local * (long long)K;
      if (prm->optimize_coefficients) {
        c->prof.points_refit += rec.n_local;
        c->prof.bytes_refit += 12ll * rec.n_local;
      }
      c->prof.points_compact += rec.n_local;
      c->prof.bytes_compact += compact_bytes(src, (size_t)rec.n_local, true, rec.n_rem_local, rec.n_inl_local, true, true);
      c->prof.points_kept += rec.n_rem_local;
      c->prof.points_peeled += rec.n_inl_local;
    }
    if (infos) infos[cur.planes] = inf;
    if (rec.t[pr::kStampEnd] != 0ull) {
      // kernel-to-kernel times of the round from the device's own stamps (stages a round does not have stay 0)
      static_assert(pr::kStampEnd == PR_LOOP_STAGES, "pr_profile.loop_ms has one entry per stage");
      unsigned long long next = rec.t[pr::kStampEnd];
      for (int i = pr::kStampEnd - 1; i >= 0; --i) {
        if (rec.t[i] == 0ull) continue;
        if (next > rec.t[i]) c->prof.loop_ms[i] += (double)(next - rec.t[i]) * 1e-6;
        next = rec.t[i];
      }
      ++c->prof.loop_rounds;
      c->timeline.insert(c->timeline.end(), rec.t, rec.t + pr::kStampSlots);
    }
    if (!rec.accepted) {
      stop_seen = true;
      *stopped = true;
      continue;
    }
    std::memcpy(coeffs + 4 * cur.planes, rec.refined, 4 * sizeof(float));
    const size_t begin = (size_t)rec.inl_off;
    cur.off = begin + (size_t)rec.n_inl_local;
    plane_offsets[cur.planes + 1] = cur.off;
    ++cur.planes;
    PR_TRY(lists.on_plane(c, begin, cur.off, stagger_lists && r + 1 < launched ? c->head_ev[r + 1] : c->round_ev[r]));
    cur.src = dst_of[r];
    cur.n_local = (size_t)rec.n_rem_local;
    cur.n_global = rec.n_rem_global;
    cur.first = c->comm ? rec.first_after : 0;
    n_bound = cur.n_local;
  }
  if (c->p2p_on) {
    PR_CUDA(cudaMemcpyAsync(c->h_p2p_err.p, c->d_p2p_aux.p + 4, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    PR_TRY(sync_stream(c));
    PR_TRY(p2p_check(c));
  }
  return PR_OK;
}

int reserve_work(plane_ransac_ctx* c) {
  const size_t cap = c->staged.cap;
  for (int i = 0; i < 2; ++i) {
    PR_TRY(dev_reserve(c->work_mem[i], 3 * cap));
    PR_TRY(dev_reserve(c->work_orig[i], cap));
    c->work[i] = planes_view(c->work_mem[i].p, c->work_orig[i].p, cap);
  }
  PR_TRY(dev_reserve(c->d_inl_cur, std::max<size_t>(c->n_staged, 1)));
  PR_TRY(dev_reserve(c->d_inl_orig, std::max<size_t>(c->n_staged, 1)));
  return PR_OK;
}

// ---- cell-ordered copy of the current cloud: shared by normal estimation and clusterFilt ----------------------------
// Uniform grid over the bounding box of the staged cloud (a superset of the current one), cell = radius (1 + 1e-6), so
// that every neighbour within the radius lies in the 27 cells around a point's cell; points sorted by cell key.
// Leaves sorted keys in d_nrm_keys[n..2n), sorted indices in d_nrm_idx[n..2n), sorted coordinates in d_nrm_xyz.
int build_cell_order(plane_ransac_ctx* c, double radius, pr::NormalsGrid* grid) {
  const size_t n = c->n_current;
  pr::NormalsGrid g;
  const double h = radius * (1.0 + 1e-6);
  g.inv_h = 1.0 / h;
  double cells = 1.0;
  for (int a = 0; a < 3; ++a) {
    double lo = 0.0, hi = 0.0;
    if (c->bbox_keys[a] <= c->bbox_keys[3 + a]) {
      lo = (double)pr::key_to_float(c->bbox_keys[a]);
      hi = (double)pr::key_to_float(c->bbox_keys[3 + a]);
    }
    g.lo[a] = lo;
    const double d = std::floor((hi - lo) * g.inv_h) + 1.0;
    if (!(d >= 1.0) || d > 2097152.0) return fail(PR_ERR_INVALID, "radius %g is too small for the cloud extent %g", radius, hi - lo);
    g.dim[a] = (long long)d;
    cells *= d;
  }
  if (cells > 9.0e18) return fail(PR_ERR_INVALID, "radius too small for the cloud extent");
  g.no_cell = (unsigned long long)g.dim[0] * (unsigned long long)g.dim[1] * (unsigned long long)g.dim[2];
  int key_bits = 1;
  while (key_bits < 64 && (g.no_cell >> key_bits) != 0) ++key_bits;
  PR_TRY(dev_reserve(c->d_nrm_keys, 2 * n));
  PR_TRY(dev_reserve(c->d_nrm_idx, 2 * n));
  PR_TRY(dev_reserve(c->d_nrm_xyz, 3 * n));
  const size_t tb = pr::normals_sort_temp_bytes(n);
  PR_TRY(dev_reserve(c->d_nrm_temp, tb + 256));
  {
    Span sp(c, KC_OTHER, 3);
    pr::launch_normals_sort(c->current, n, g, key_bits, c->d_nrm_keys.p, c->d_nrm_idx.p, c->d_nrm_temp.p, tb, c->d_nrm_xyz.p, c->stream);
  }
  *grid = g;
  return PR_OK;
}
// Maps every peer's mailbox (CUDA IPC over NVLink).  Collective; all ranks end with the same p2p_on: if any rank
// cannot map a peer (no P2P path, IPC unavailable) or PR_P2P=0 is set, every rank keeps the NCCL exchanges.
int p2p_setup(plane_ransac_ctx* c) {
  c->p2p_on = false;
  if (c->n_ranks < 2) return PR_OK;
  const char* env = getenv("PR_P2P");
  int ok = (c->n_ranks <= pr::kP2PMaxRanks && !(env && atoi(env) == 0)) ? 1 : 0;
  PR_TRY(dev_reserve(c->d_p2p_aux, 8));
  PR_TRY(pin_reserve(c->h_p2p_err, 1));
  *c->h_p2p_err.p = 0;
  PR_CUDA(cudaMemsetAsync(c->d_p2p_aux.p, 0, 8 * sizeof(unsigned), c->stream));
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok) {
    if (cudaMalloc(&c->p2p_mailbox, kP2PMailboxBytes) != cudaSuccess || cudaMemset(c->p2p_mailbox, 0, kP2PMailboxBytes) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess || cudaIpcGetMemHandle(&mine, c->p2p_mailbox) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
    }
  }
  // all-gather of the handles (64 bytes each) and of the per-rank verdicts through NCCL
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  const size_t rec = 64 + 8;
  DevBuf<unsigned char> d_h;
  PR_TRY(dev_reserve(d_h, rec * (size_t)(c->n_ranks + 1)));
  std::vector<unsigned char> h((size_t)(c->n_ranks + 1) * rec, 0);
  std::memcpy(h.data(), &mine, 64);
  h[64] = (unsigned char)ok;
  int rc = PR_OK;
  auto gather = [&]() -> int {
    PR_CUDA(cudaMemcpyAsync(d_h.p, h.data(), rec, cudaMemcpyHostToDevice, c->stream));
    PR_NCCL(g_nccl.AllGather(d_h.p, d_h.p + rec, rec, ncclUint8, c->comm, c->stream));
    PR_CUDA(cudaMemcpyAsync(h.data() + rec, d_h.p + rec, rec * (size_t)c->n_ranks, cudaMemcpyDeviceToHost, c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
    return PR_OK;
  };
  rc = gather();
  if (rc == PR_OK) {
    for (int r = 0; r < c->n_ranks; ++r) ok = ok && h[rec * (size_t)(r + 1) + 64];
    c->p2p_view.n_ranks = c->n_ranks;
    c->p2p_view.rank = c->rank;
    for (int r = 0; r < pr::kP2PMaxRanks; ++r) c->p2p_view.peers[r] = nullptr;
    if (ok) {
      for (int r = 0; r < c->n_ranks; ++r) {
        if (r == c->rank) {
          c->p2p_view.peers[r] = c->p2p_mailbox;
          continue;
        }
        cudaIpcMemHandle_t hr;
        std::memcpy(&hr, h.data() + rec * (size_t)(r + 1), 64);
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = 0;
          break;
        }
        c->p2p_view.peers[r] = static_cast<unsigned char*>(ptr);
      }
    }
    // second round: did every rank map every peer?
    h[64] = (unsigned char)ok;
    rc = gather();
    if (rc == PR_OK)
      for (int r = 0; r < c->n_ranks; ++r) ok = ok && h[rec * (size_t)(r + 1) + 64];
  }
  dev_free(d_h);
  if (!ok) p2p_teardown(c, rc == PR_OK);
  if (rc != PR_OK) ok = 0;
  c->p2p_on = ok != 0;
  if (rc == PR_OK) {
    PR_TRY(dev_reserve(c->d_p2p_epoch, 8));  // 4 epochs, then the tickets of the multi-block exchanges (2 x uint32 each)
    PR_TRY(dev_reserve(c->d_p2p_wait, 8));
    PR_CUDA(cudaMemsetAsync(c->d_p2p_epoch.p, 0, 8 * sizeof(unsigned long long), c->stream));
    PR_CUDA(cudaMemsetAsync(c->d_p2p_wait.p, 0, 8 * sizeof(unsigned long long), c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
  }
  return rc;
}

}  // namespace

namespace pr {
// Pageable stand-in for plane_ransac_host_alloc where page-locking is impossible (no CUDA device: the PCD reader still
// works there, e.g. in a conversion tool); remembered so that plane_ransac_host_free releases it the right way.
static std::mutex g_plain_mu;
static std::vector<void*> g_plain_allocs;
void* plain_alloc(size_t bytes) {
  void* p = std::malloc(bytes ? bytes : 1);
  if (p) {
    std::lock_guard<std::mutex> lk(g_plain_mu);
    g_plain_allocs.push_back(p);
  }
  return p;
}
bool release_plain_alloc(void* p) {
  std::lock_guard<std::mutex> lk(g_plain_mu);
  for (size_t i = 0; i < g_plain_allocs.size(); ++i)
    if (g_plain_allocs[i] == p) {
      g_plain_allocs.erase(g_plain_allocs.begin() + (long)i);
      std::free(p);
      return true;
    }
  return false;
}
// error hook for the other translation units of the library (pr_pcd.cpp)
int pcd_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}
}  // namespace pr

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

int plane_ransac_abi_version(void) { return PLANE_RANSAC_ABI_VERSION; }

const char* plane_ransac_last_error(void) { return g_error.c_str(); }

void plane_ransac_default_params(pr_params* p) {
  if (!p) return;
  p->distance_threshold = 0.1;  // Dialog/config.txt:29 T_dist_point_plane
  p->max_iterations = 50;       // pcl::SACSegmentation default
  p->min_plane_size = 500;      // Dialog/config.txt:20 T_num_of_single_plane
  p->probability = 0.99;
  p->optimize_coefficients = 1;
  p->seed = 12345u;
  p->max_planes = 64;
  p->dot_order = PR_DOT_FMA;
  p->scorer = PR_SCORER_BRUTE;
  p->refit_mode = PR_REFIT_FIXED;
}

int plane_ransac_create(plane_ransac_ctx** out, int device_id) {
  if (!out) return fail(PR_ERR_INVALID, "null output pointer");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(PR_ERR_CUDA, "no CUDA device available (%s); this backend has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device_id < 0 || device_id >= n_dev) return fail(PR_ERR_INVALID, "device_id %d out of range (0..%d)", device_id, n_dev - 1);
  PR_CUDA(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  PR_CUDA(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major < 10) return fail(PR_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device_id, prop.major, prop.minor);
  plane_ransac_ctx* c = new (std::nothrow) plane_ransac_ctx();
  if (!c) return fail(PR_ERR_OOM, "out of host memory");
  c->device = device_id;
  c->num_sms = prop.multiProcessorCount;
  std::memset(&c->prof, 0, sizeof(c->prof));
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->copy_ready, cudaEventDisableTiming) != cudaSuccess) {
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
    return fail(PR_ERR_CUDA, "cudaStreamCreate failed");
  }
  *out = c;
  return PR_OK;
}

void plane_ransac_destroy(plane_ransac_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  collect_spans(c);
  p2p_teardown(c, c->p2p_on);
  dev_free(c->d_p2p_aux);
  dev_free(c->d_p2p_epoch);
  dev_free(c->d_p2p_wait);
  pin_free(c->h_p2p_err);
  dev_free(c->d_state); pin_free(c->h_state); pin_free(c->h_recs); dev_free(c->d_chain_tickets);
  dev_free(c->d_rnd); dev_free(c->d_draw_table); dev_free(c->d_draw_coll);
  for (cudaEvent_t e : c->round_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : c->head_ev) cudaEventDestroy(e);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  dev_free(c->staged_mem);
  for (int i = 0; i < 2; ++i) { dev_free(c->work_mem[i]); dev_free(c->work_orig[i]); }
  dev_free(c->aos); dev_free(c->d_bbox); dev_free(c->d_triples); dev_free(c->d_counts); dev_free(c->d_good);
  dev_free(c->d_sample_pts); dev_free(c->d_hyps); dev_free(c->d_refit); dev_free(c->d_totals);
  dev_free(c->d_scratch); dev_free(c->d_pcl_sums); dev_free(c->d_inl_cur); dev_free(c->d_inl_orig); dev_free(c->d_flush); dev_free(c->batch_mem);
  dev_free(c->d_stage_map);
  dev_free(c->d_stage_map_tmp);
  for (int i = 0; i < 3; ++i) dev_free(c->sorted_mem[i]);
  dev_free(c->d_bounds); dev_free(c->d_aux); dev_free(c->d_keys); dev_free(c->d_vals); dev_free(c->d_sort_temp);
  dev_free(c->d_totals2);
  dev_free(c->d_rb_planes); dev_free(c->d_rb_border); dev_free(c->d_rb_edges); dev_free(c->d_rb_rays);
  dev_free(c->d_rb_ray_edges); dev_free(c->d_rb_counts); dev_free(c->d_rb_out_cur); dev_free(c->d_rb_out_orig);
  dev_free(c->d_rb_counters); dev_free(c->d_rb_keys); dev_free(c->d_rb_cand); dev_free(c->d_rb_claimed); dev_free(c->d_rb_temp);
  dev_free(c->d_nrm_keys); dev_free(c->d_nrm_idx); dev_free(c->d_nrm_xyz); dev_free(c->d_nrm_out); dev_free(c->d_nrm_cnt);
  dev_free(c->d_nrm_temp);
  dev_free(c->d_batch_bbox); dev_free(c->d_batch_idx); dev_free(c->d_batch_scale); dev_free(c->d_batch_refit);
  dev_free(c->d_batch_hyps); dev_free(c->d_batch_pts); dev_free(c->d_batch_cnt);
  dev_free(c->d_batch_sexp); dev_free(c->d_batch_lists);
  dev_free(c->d_batch_offs);
  dev_free(c->d_batch_out); pin_free(c->h_batch_out); dev_free(c->d_batch_tri);
  pin_free(c->h_triples); pin_free(c->h_counts); pin_free(c->h_good); pin_free(c->h_refit); pin_free(c->h_sample_pts);
  for (cudaEvent_t e : c->pend.ev) cudaEventDestroy(e);
  pin_free(c->h_totals); pin_free(c->h_small);
  if (c->timer_a) { cudaEventDestroy(c->timer_a); cudaEventDestroy(c->timer_b); }
  for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
  cudaStreamSynchronize(c->copy_stream);
  cudaEventDestroy(c->copy_ready);
  cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->stream);
  delete c;
}

int plane_ransac_set_cloud(plane_ransac_ctx* c, const pr_point* pts, size_t n) {
  PR_TRY(check_ctx(c, true));
  cancel_pending(c);
  if (!pts && n) return fail(PR_ERR_INVALID, "null cloud");
  if (n > (size_t)INT_MAX - 4096) return fail(PR_ERR_INVALID, "cloud too large for 32-bit indices");
  PR_TRY(dev_reserve(c->aos, std::max<size_t>(n, 1)));
  if (n) PR_CUDA(cudaMemcpyAsync(c->aos.p, pts, n * sizeof(pr_point), cudaMemcpyHostToDevice, c->stream));
  return stage_from_device(c, c->aos.p, n);
}

int plane_ransac_set_cloud_async(plane_ransac_ctx* c, const pr_point* pts, size_t n) {
  PR_TRY(check_ctx(c, true));
  cancel_pending(c);
  if (!pts && n) return fail(PR_ERR_INVALID, "null cloud");
  if (n > (size_t)INT_MAX - 4096) return fail(PR_ERR_INVALID, "cloud too large for 32-bit indices");
  // small clouds and pageable memory (the copy would block the host anyway) take the synchronous path; every rank of
  // a sharded context must therefore make the same choice (same size class, pinned everywhere)
  cudaPointerAttributes attr;
  const bool pinned = n && cudaPointerGetAttributes(&attr, pts) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!pinned || n < ((size_t)1 << 21)) return plane_ransac_set_cloud(c, pts, n);

  const size_t cap = pr::padded_capacity(n);
  PR_TRY(dev_reserve(c->aos, n));
  PR_TRY(dev_reserve(c->staged_mem, 3 * cap));
  PR_TRY(reserve_small(c));
  c->staged = planes_view(c->staged_mem.p, nullptr, cap);
  c->have_stage_map = false;
  c->last_offsets.clear();
  c->last_coeffs.clear();
  c->sorted_staged_valid = false;
  auto& u = c->pend;
  // Chunk sizes grow, then shrink.  The first scoring launch can start once the first chunk has landed, so that one is
  // small.  While scoring a chunk takes longer than copying it (one GPU: ~5.4 ms against ~3 ms for 10M points x 4096
  // hypotheses) the device never waits for the copy again and larger chunks score more efficiently; when the copy is
  // the slower one (eight ranks sharing the host's memory system) the call ends one chunk's scoring after the copy, so
  // the last chunks are small again.  PR_UPLOAD_WEIGHTS=1,1,1,1,1,1,1,1 gives round 1's eight equal chunks.
  static const std::vector<int> weights = [] {
    std::vector<int> w;
    if (const char* e = getenv("PR_UPLOAD_WEIGHTS")) {
      for (const char* p = e; *p;) {
        char* end = nullptr;
        const long v = strtol(p, &end, 10);
        if (end == p) break;
        if (v > 0) w.push_back((int)v);
        p = *end ? end + 1 : end;
      }
    }
    if (w.empty()) w = {1, 2, 3, 4, 3, 2, 1};
    return w;
  }();
  {
    long long wsum = 0;
    for (int w : weights) wsum += w;
    const size_t tiles = (n + pr::kTilePoints - 1) / pr::kTilePoints;
    u.off.assign(1, 0);
    long long acc = 0;
    for (size_t k = 0; k < weights.size(); ++k) {
      acc += weights[k];
      const size_t end = k + 1 == weights.size() ? n : std::min(n, (size_t)((double)tiles * (double)acc / (double)wsum) * (size_t)pr::kTilePoints);
      if (end > u.off.back()) u.off.push_back(end);
    }
    u.n_chunks = (int)u.off.size() - 1;
  }
  while ((int)u.ev.size() < u.n_chunks) {
    cudaEvent_t e = nullptr;
    PR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    u.ev.push_back(e);
  }
  pr::launch_bbox_init(c->d_bbox.p, c->stream);
  c->prof.launches_stage += 1;
  for (int k = 0; k < u.n_chunks; ++k) {
    const size_t off = u.off[k], cnt = u.off[k + 1] - off;
    PR_CUDA(cudaMemcpyAsync(c->aos.p + off, pts + off, cnt * sizeof(pr_point), cudaMemcpyHostToDevice, c->copy_stream));
    PR_CUDA(cudaEventRecord(u.ev[k], c->copy_stream));
  }
  u.active = true;
  u.host = pts;
  u.n = n;
  u.next = 0;
  c->n_staged = n;
  c->have_cloud = true;
  c->current = c->staged;
  c->n_current = n;
  c->n_global_staged = c->n_global_current = (long long)n;
  c->first_staged = c->first_current = 0;
  c->global_valid = false;
  if (c->comm) {
    // the global size and this rank's offset are needed before the first draw; the bounding box follows when the
    // last chunk is staged (finalize_pending)
    long long* d = c->d_totals.p;
    c->h_totals.p[0] = (long long)n;
    PR_CUDA(cudaMemcpyAsync(d, c->h_totals.p, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    PR_NCCL(g_nccl.AllGather(d, d + 2, 1, ncclInt64, c->comm, c->stream));
    PR_CUDA(cudaMemcpyAsync(c->h_totals.p + 2, d + 2, c->n_ranks * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
    long long first = 0, total = 0;
    for (int r = 0; r < c->n_ranks; ++r) {
      if (r == c->rank) first = total;
      total += c->h_totals.p[2 + r];
    }
    c->n_global_staged = c->n_global_current = total;
    c->first_staged = c->first_current = first;
  }
  return PR_OK;
}

int plane_ransac_set_cloud_ex(plane_ransac_ctx* c, const pr_point* pts, size_t n, unsigned flags, size_t* n_kept,
                              float centroid[3]) {
  PR_TRY(check_ctx(c, true));
  cancel_pending(c);
  if (!pts && n) return fail(PR_ERR_INVALID, "null cloud");
  if (n > (size_t)INT_MAX - 4096) return fail(PR_ERR_INVALID, "cloud too large for 32-bit indices");
  if (flags & ~(unsigned)(PR_STAGE_REMOVE_NONFINITE | PR_STAGE_TRANSLATE_CENTROID)) return fail(PR_ERR_INVALID, "unknown staging flag");
  PR_TRY(dev_reserve(c->aos, std::max<size_t>(n, 1)));
  if (n) PR_CUDA(cudaMemcpyAsync(c->aos.p, pts, n * sizeof(pr_point), cudaMemcpyHostToDevice, c->stream));
  return stage_from_device(c, c->aos.p, n, flags, n_kept, centroid);
}

int plane_ransac_staged_source_indices(plane_ransac_ctx* c, int32_t* out, size_t cap) {
  PR_TRY(check_ctx(c));
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (c->n_staged > cap || (!out && c->n_staged)) return fail(PR_ERR_CAPACITY, "output holds %zu indices, need %zu", cap, c->n_staged);
  if (c->n_staged == 0) return PR_OK;
  if (c->have_stage_map) {
    PR_CUDA(cudaMemcpyAsync(out, c->d_stage_map.p, c->n_staged * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
  } else {
    for (size_t i = 0; i < c->n_staged; ++i) out[i] = (int32_t)i;
  }
  return PR_OK;
}

int plane_ransac_set_cloud_device(plane_ransac_ctx* c, const pr_point* dev_pts, size_t n) {
  PR_TRY(check_ctx(c, true));
  cancel_pending(c);
  if (!dev_pts && n) return fail(PR_ERR_INVALID, "null cloud");
  if (n > (size_t)INT_MAX - 4096) return fail(PR_ERR_INVALID, "cloud too large for 32-bit indices");
  cudaPointerAttributes attr;
  if (n && (cudaPointerGetAttributes(&attr, dev_pts) != cudaSuccess || attr.type != cudaMemoryTypeDevice)) {
    cudaGetLastError();
    return fail(PR_ERR_INVALID, "set_cloud_device needs a device pointer");
  }
  return stage_from_device(c, reinterpret_cast<const float4*>(dev_pts), n);
}

int plane_ransac_cloud_size(plane_ransac_ctx* c, size_t* n_staged, size_t* n_current) {
  if (!c) return fail(PR_ERR_INVALID, "null context");
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (n_staged) *n_staged = c->n_staged;
  if (n_current) *n_current = c->n_current;
  return PR_OK;
}

int plane_ransac_score(plane_ransac_ctx* c, const int32_t* triples, int K, double t, int dot_order, int32_t* counts,
                       float* coeffs, uint8_t* good) {
  PR_TRY(check_ctx(c));
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (K < 0 || (K && (!triples || !counts))) return fail(PR_ERR_INVALID, "bad triples/counts");
  if (!(t > 0.0)) return fail(PR_ERR_INVALID, "threshold must be > 0");
  if (dot_order != PR_DOT_PCL_SSE2 && dot_order != PR_DOT_FMA) return fail(PR_ERR_INVALID, "unknown dot_order");
  if (K == 0) return PR_OK;
  for (int i = 0; i < 3 * K; ++i)
    if (triples[i] < 0 || (long long)triples[i] >= c->n_global_staged) return fail(PR_ERR_INVALID, "triple index %d out of range", triples[i]);
  PR_TRY(reserve_draws(c, (size_t)K, false));
  std::memcpy(c->h_triples.p, triples, 3 * (size_t)K * sizeof(int32_t));
  PR_CUDA(cudaMemcpyAsync(c->d_triples.p, c->h_triples.p, 3 * (size_t)K * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  {
    Span sp(c, KC_MODELS, 2);
    pr::launch_gather_samples(c->staged, c->first_staged, c->n_staged, c->d_triples.p, 3 * K, c->d_sample_pts.p, 1, 0, c->stream);
    if (c->comm) PR_NCCL(g_nccl.AllReduce(c->d_sample_pts.p, c->d_sample_pts.p, (size_t)(3 * K) * 4, ncclInt32, ncclSum, c->comm, c->stream));
    pr::launch_models(c->d_sample_pts.p, K, c->d_hyps.p, c->d_good.p, c->stream);
  }
  PR_CUDA(cudaMemsetAsync(c->d_counts.p, 0, (size_t)K * sizeof(int32_t), c->stream));
  {
    Span sp(c, KC_SCORE, 0);
    c->prof.launches_score += pr::launch_score(c->staged, c->n_staged, 1, 0, c->d_hyps.p, K, pr::threshold_up(t), dot_order, c->d_counts.p, c->num_sms, c->stream);
    c->prof.pairs_scored += (long long)c->n_staged * K;
  }
  if (c->comm) PR_NCCL(g_nccl.AllReduce(c->d_counts.p, c->d_counts.p, (size_t)K, ncclInt32, ncclSum, c->comm, c->stream));
  PR_CUDA(cudaGetLastError());
  PR_CUDA(cudaMemcpyAsync(c->h_counts.p, c->d_counts.p, (size_t)K * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  PR_CUDA(cudaMemcpyAsync(c->h_good.p, c->d_good.p, (size_t)K * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (coeffs) PR_CUDA(cudaMemcpyAsync(coeffs, c->d_hyps.p, (size_t)K * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  PR_CUDA(cudaStreamSynchronize(c->stream));
  std::memcpy(counts, c->h_counts.p, (size_t)K * sizeof(int32_t));
  if (good)
    for (int k = 0; k < K; ++k) good[k] = c->h_good.p[k] ? 1 : 0;
  return PR_OK;
}

int plane_ransac_segment_one(plane_ransac_ctx* c, const pr_params* prm, float coeff[4], int32_t* inliers, size_t cap,
                             size_t* n_inliers, pr_segment_info* info) {
  PR_TRY(check_ctx(c, true));
  if (c->profiling) collect_spans(c);  // stream is idle here: fold finished spans, recycle their events
  PR_TRY(check_params(prm));
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (!coeff || !n_inliers) return fail(PR_ERR_INVALID, "null output");
  PR_TRY(dev_reserve(c->d_inl_cur, std::max<size_t>(c->n_staged, 1)));
  SegmentOut so;
  pr::CloudView none;
  const bool hier = prm->scorer == PR_SCORER_HIER;
  if (hier) PR_TRY(ensure_staged(c));
  if (hier) PR_TRY(ensure_sorted(c, 1));
  PR_TRY(segment_core(c, prm, c->staged, c->n_staged, c->n_global_staged, c->first_staged, false, none, c->d_inl_cur.p,
                      nullptr, info, &so, hier ? &c->sorted_view[0] : nullptr, nullptr));
  if (c->pend.active) PR_TRY(ensure_staged(c));
  std::memcpy(coeff, so.coeff, 4 * sizeof(float));
  *n_inliers = (size_t)so.n_inl_local;
  if (inliers) {
    if ((size_t)so.n_inl_local > cap) return fail(PR_ERR_CAPACITY, "inlier buffer holds %zu, need %lld", cap, so.n_inl_local);
    if (so.n_inl_local) {
      PR_CUDA(cudaMemcpyAsync(inliers, c->d_inl_cur.p, (size_t)so.n_inl_local * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
      PR_CUDA(cudaStreamSynchronize(c->stream));
    }
  }
  return PR_OK;
}

int plane_ransac_extract_planes(plane_ransac_ctx* c, const pr_params* prm, float* coeffs, int32_t* inlier_cur,
                                int32_t* inlier_orig, size_t idx_cap, size_t* plane_offsets, int* n_planes,
                                pr_segment_info* infos) {
  PR_TRY(check_ctx(c, true));
  if (c->profiling) collect_spans(c);  // stream is idle here: fold finished spans, recycle their events
  PR_TRY(check_params(prm));
  PR_TRY(check_comm(c));
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (!coeffs || !plane_offsets || !n_planes) return fail(PR_ERR_INVALID, "null output");
  PR_TRY(reserve_work(c));
  const bool hier = prm->scorer == PR_SCORER_HIER;
  if (hier) PR_TRY(ensure_staged(c));
  if (hier) PR_TRY(ensure_sorted(c, 3));
  PeelCursor cur;
  cur.src = c->staged;
  cur.n_local = c->n_staged;
  cur.n_global = c->n_global_staged;
  cur.first = c->first_staged;
  plane_offsets[0] = 0;
  *n_planes = 0;
  ListCopier lists;
  lists.cur = inlier_cur;
  lists.orig = inlier_orig;
  lists.cap = idx_cap;
  // Pinned destination buffers: each plane's index lists go back on the copy stream as soon as its round is done,
  // under the scoring of the following rounds (a pageable destination would make the copy block the host instead).
  auto is_pinned = [](const void* p) {
    cudaPointerAttributes a;
    if (!p) return true;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  lists.overlap = lists.wanted() && is_pinned(inlier_cur) && is_pinned(inlier_orig);
  // Any failure below leaves the context on the staged cloud (work[0] / work[1] may hold a half-finished round) with
  // no copy still in flight towards the caller's buffers.
  auto bail = [&](int rc) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->stream);
    c->current = c->staged;
    c->n_current = c->n_staged;
    c->n_global_current = c->n_global_staged;
    c->first_current = c->first_staged;
    c->last_offsets.clear();
    c->last_coeffs.clear();
    return rc;
  };
  int handed_back = 0;
  bool finished = false;
  c->timeline.clear();
  while (!finished && cur.planes < prm->max_planes) {
    if (chain_eligible(c, prm, cur.n_global) && handed_back < 2) {
      bool stopped = false, back = false;
      const int rc = run_chain(c, prm, cur, coeffs, plane_offsets, infos, lists, &stopped, &back);
      if (rc != PR_OK) return bail(rc);
      if (stopped) break;
      if (!back) continue;  // every queued round was accepted: the loop condition ends the call
      ++handed_back;        // the round in flight needs the host (crowded sampler or a bad sample): one host-driven round
    }
    pr::CloudView dst = c->work[cur.planes & 1];
    SegmentOut so;
    pr_segment_info inf;
    const int s_next = 1 + (cur.planes & 1);
    const int rc = segment_core(c, prm, cur.src, cur.n_local, cur.n_global, cur.first, true, dst, c->d_inl_cur.p + cur.off,
                                c->d_inl_orig.p + cur.off, &inf, &so, hier ? &c->sorted_view[cur.s_cur] : nullptr,
                                hier ? &c->sorted_view[s_next] : nullptr);
    if (rc != PR_OK) return bail(rc);
    if (infos) infos[cur.planes] = inf;
    const long long m = so.n_inl_global;
    if (m == 0 || m < (long long)std::max(0, prm->min_plane_size)) break;
    std::memcpy(coeffs + 4 * cur.planes, so.coeff, 4 * sizeof(float));
    const size_t begin = cur.off;
    cur.off += (size_t)so.n_inl_local;
    plane_offsets[cur.planes + 1] = cur.off;
    ++cur.planes;
    // segment_core ended with a stream synchronisation: the lists are complete
    { const int rc2 = lists.on_plane(c, begin, cur.off, nullptr); if (rc2 != PR_OK) return bail(rc2); }
    cur.src = dst;
    cur.s_cur = s_next;
    cur.n_local = (size_t)so.n_rem_local;
    if (c->comm) {
      long long tot = 0, f = 0;
      for (int r = 0; r < c->n_ranks; ++r) {
        if (r == c->rank) f = tot;
        tot += so.rem_per_rank[r];
      }
      cur.n_global = tot;
      cur.first = f;
    } else {
      cur.n_global = (long long)cur.n_local;
      cur.first = 0;
    }
  }
  if (c->pend.active) { const int rc = ensure_staged(c); if (rc != PR_OK) return bail(rc); }
  *n_planes = cur.planes;
  c->last_offsets.assign(plane_offsets, plane_offsets + cur.planes + 1);
  c->last_coeffs.assign(coeffs, coeffs + 4 * (size_t)cur.planes);
  c->current = cur.src;
  c->n_current = cur.n_local;
  c->n_global_current = cur.n_global;
  c->first_current = cur.first;
  if (lists.overlap) {
    PR_CUDA(cudaStreamSynchronize(c->copy_stream));
  } else if (!lists.overflow) {
    if (cur.off && inlier_cur) PR_CUDA(cudaMemcpyAsync(inlier_cur, c->d_inl_cur.p, cur.off * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (cur.off && inlier_orig) PR_CUDA(cudaMemcpyAsync(inlier_orig, c->d_inl_orig.p, cur.off * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  PR_TRY(sync_stream(c));
  // The peel itself is complete and consistent (every rank ran every collective); only the caller's index buffers were
  // too small: coefficients, offsets and plane_ransac_remaining / plane_points are valid, the lists are not.
  if (lists.overflow)
    return fail(PR_ERR_CAPACITY, "inlier index buffers hold %zu entries, need at least %zu", idx_cap, lists.needed);
  return PR_OK;
}

int plane_ransac_set_round_loop(plane_ransac_ctx* c, int mode) {
  if (!c) return fail(PR_ERR_INVALID, "null context");
  if (mode != PR_LOOP_AUTO && mode != PR_LOOP_HOST) return fail(PR_ERR_INVALID, "unknown round-loop mode");
  c->round_loop = mode;
  return PR_OK;
}

int plane_ransac_plane_points(plane_ransac_ctx* c, int k, int project, pr_point* out, size_t cap, size_t* n) {
  PR_TRY(check_ctx(c));
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (k < 0 || (size_t)k + 1 >= c->last_offsets.size()) return fail(PR_ERR_INVALID, "plane %d is not a plane of the last extract call", k);
  const size_t off = c->last_offsets[k], cnt = c->last_offsets[k + 1] - off;
  if (n) *n = cnt;
  if (!out) return PR_OK;
  if (cnt > cap) return fail(PR_ERR_CAPACITY, "output holds %zu points, need %zu", cap, cnt);
  if (cnt == 0) return PR_OK;
  PR_TRY(dev_reserve(c->aos, std::max(cnt, c->aos.cap)));
  {
    Span sp(c, KC_OTHER, 1);
    pr::Plane4 pl = {c->last_coeffs[4 * k], c->last_coeffs[4 * k + 1], c->last_coeffs[4 * k + 2], c->last_coeffs[4 * k + 3]};
    pr::launch_plane_points(c->staged, c->d_inl_orig.p + off, cnt, pl, project != 0, c->aos.p, c->stream);
  }
  PR_CUDA(cudaGetLastError());
  PR_CUDA(cudaMemcpyAsync(out, c->aos.p, cnt * sizeof(pr_point), cudaMemcpyDeviceToHost, c->stream));
  PR_CUDA(cudaStreamSynchronize(c->stream));
  return PR_OK;
}

int plane_ransac_remaining(plane_ransac_ctx* c, pr_point* out, size_t cap, size_t* n) {
  PR_TRY(check_ctx(c));
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (n) *n = c->n_current;
  if (!out) return PR_OK;
  if (c->n_current > cap) return fail(PR_ERR_CAPACITY, "output holds %zu points, need %zu", cap, c->n_current);
  if (c->n_current == 0) return PR_OK;
  PR_TRY(dev_reserve(c->aos, c->n_current));
  {
    Span sp(c, KC_STAGE, 1);
    pr::launch_unstage(c->current, c->n_current, c->aos.p, c->stream);
  }
  PR_CUDA(cudaGetLastError());
  PR_CUDA(cudaMemcpyAsync(out, c->aos.p, c->n_current * sizeof(pr_point), cudaMemcpyDeviceToHost, c->stream));
  PR_CUDA(cudaStreamSynchronize(c->stream));
  return PR_OK;
}


// ---- pcl::NormalEstimationOMP with a radius search (Dialog/PlaneDetect.h:515-545) ----------------------------------
int plane_ransac_estimate_normals(plane_ransac_ctx* c, double radius, const float viewpoint[3], pr_normal* out, size_t cap,
                                  int32_t* n_neighbors) {
  PR_TRY(check_ctx(c));
  if (c->profiling) collect_spans(c);
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (!(radius > 0.0) || !std::isfinite(radius)) return fail(PR_ERR_INVALID, "radius must be finite and > 0");
  if (c->comm) return fail(PR_ERR_INVALID, "normal estimation needs the whole cloud on one GPU (neighbourhoods cross shard boundaries)");
  const size_t n = c->n_current;
  if (n > cap || (!out && n)) return fail(PR_ERR_CAPACITY, "output holds %zu normals, need %zu", cap, n);
  if (n == 0) return PR_OK;
  pr::NormalsGrid g;
  PR_TRY(build_cell_order(c, radius, &g));
  int e = 0;
  (void)std::frexp(radius, &e);  // radius < 2^e: neighbour offsets times 2^(18 - e) stay below 2^18
  const double scale = std::ldexp(1.0, 18 - e);
  const float r2 = (float)(radius * radius);
  const float vp[3] = {viewpoint ? viewpoint[0] : 0.f, viewpoint ? viewpoint[1] : 0.f, viewpoint ? viewpoint[2] : 0.f};
  PR_TRY(dev_reserve(c->d_nrm_out, n));
  PR_TRY(dev_reserve(c->d_nrm_cnt, n));
  {
    Span sp(c, KC_OTHER, 1);
    pr::launch_normals(c->d_nrm_xyz.p, c->d_nrm_keys.p + n, c->d_nrm_idx.p + n, n, g, r2, scale, vp, c->d_nrm_out.p,
                       n_neighbors ? c->d_nrm_cnt.p : nullptr, c->stream);
  }
  PR_CUDA(cudaGetLastError());
  PR_CUDA(cudaMemcpyAsync(out, c->d_nrm_out.p, n * sizeof(pr_normal), cudaMemcpyDeviceToHost, c->stream));
  if (n_neighbors) PR_CUDA(cudaMemcpyAsync(n_neighbors, c->d_nrm_cnt.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  PR_TRY(sync_stream(c));
  return PR_OK;
}

// ---- clusterFilt (Dialog/PlaneDetect.h:1582-1656) --------------------------------------------------------------------
int plane_ransac_cluster_filter(plane_ransac_ctx* c, double radius, int max_small_cluster, size_t* n_removed, size_t* n_remaining) {
  PR_TRY(check_ctx(c));
  if (c->profiling) collect_spans(c);
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (!(radius > 0.0) || !std::isfinite(radius)) return fail(PR_ERR_INVALID, "radius must be finite and > 0");
  if (max_small_cluster < 0) return fail(PR_ERR_INVALID, "max_small_cluster must be >= 0");
  if (c->comm) return fail(PR_ERR_INVALID, "clusterFilt needs the whole cloud on one GPU (clusters cross shard boundaries)");
  const size_t n = c->n_current;
  if (n_removed) *n_removed = 0;
  if (n_remaining) *n_remaining = n;
  if (n == 0) return PR_OK;
  PR_TRY(reserve_small(c));
  PR_TRY(reserve_work(c));
  pr::NormalsGrid g;
  PR_TRY(build_cell_order(c, radius, &g));
  // union-find scratch: parent in the unsorted half of the index buffer, sizes in the count buffer
  PR_TRY(dev_reserve(c->d_nrm_cnt, n));
  PR_TRY(dev_reserve(c->d_rb_claimed, c->current.cap));
  PR_CUDA(cudaMemsetAsync(c->d_rb_claimed.p, 0, c->current.cap * sizeof(uint32_t), c->stream));
  {
    Span sp(c, KC_OTHER, 4);
    pr::launch_cluster_flags(c->d_nrm_xyz.p, c->d_nrm_keys.p + n, c->d_nrm_idx.p + n, n, g, (float)(radius * radius),
                             (uint32_t)max_small_cluster, c->d_nrm_idx.p, reinterpret_cast<uint32_t*>(c->d_nrm_cnt.p),
                             c->d_rb_claimed.p, c->stream);
  }
  PR_CUDA(cudaGetLastError());
  pr::CloudView dst = (c->current.x == c->work[0].x) ? c->work[1] : c->work[0];
  const pr::CloudView src_before = c->current;
  PR_TRY(dev_reserve(c->d_scratch, pr::compact_scratch_bytes(n) + 64));
  {
    Span sp(c, KC_COMPACT, 1);
    pr::Plane4 none = {0, 0, 0, 0};
    pr::launch_compact(c->current, n, none, 0.f, 3, dst, true, nullptr, nullptr, c->d_scratch.p, c->d_totals.p, c->stream, c->d_rb_claimed.p);
    c->prof.points_compact += (long long)n;
  }
  PR_CUDA(cudaGetLastError());
  PR_CUDA(cudaMemcpyAsync(c->h_totals.p, c->d_totals.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  PR_TRY(sync_stream(c));
  c->current = dst;
  c->n_current = (size_t)c->h_totals.p[0];
  c->n_global_current = (long long)c->n_current;
  c->first_current = 0;
  c->prof.bytes_compact += 4ll * (long long)n + compact_bytes(src_before, n, true, (long long)c->n_current, 0, false, false);
  if (n_removed) *n_removed = (size_t)c->h_totals.p[1];
  if (n_remaining) *n_remaining = c->n_current;
  return PR_OK;
}

// ---- "run again" (Dialog/PCLViewer.cpp:1120-1178): the cloud left by the last call becomes the staged cloud --------
int plane_ransac_restage_remaining(plane_ransac_ctx* c) {
  PR_TRY(check_ctx(c));
  if (c->profiling) collect_spans(c);
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  c->last_offsets.clear();
  c->last_coeffs.clear();
  if (c->current.x == c->staged.x) return PR_OK;  // nothing was peeled
  const size_t n = c->n_current, cap = c->staged.cap;
  const size_t padded = std::min(cap, pr::padded_capacity(n));  // the peel re-padded [n, padded) with NaN
  // staged point -> caller's index: compose the peel's original-index plane with the map of a filtered staging
  PR_TRY(dev_reserve(c->d_stage_map_tmp, std::max<size_t>(cap, 1)));
  {
    Span sp(c, KC_STAGE, n ? 1 : 0);
    pr::launch_compose_map(c->current.orig, c->have_stage_map ? c->d_stage_map.p : nullptr, n, c->d_stage_map_tmp.p, c->stream);
  }
  PR_CUDA(cudaGetLastError());
  std::swap(c->d_stage_map, c->d_stage_map_tmp);
  c->have_stage_map = true;
  const float* src[3] = {c->current.x, c->current.y, c->current.z};
  float* dst[3] = {c->staged.x, c->staged.y, c->staged.z};
  for (int a = 0; a < 3; ++a) PR_CUDA(cudaMemcpyAsync(dst[a], src[a], padded * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  PR_CUDA(cudaStreamSynchronize(c->stream));
  // the old bounding box still contains every point: the refit grid (scale_exp) stays valid
  c->n_staged = n;
  c->current = c->staged;
  c->n_global_staged = c->n_global_current;
  c->first_staged = c->first_current;
  c->sorted_staged_valid = false;
  return PR_OK;
}

// ---- postProcessPlanes re-absorption (Dialog/PlaneDetect.h:1454-1580) ------------------------------
int plane_ransac_reabsorb(plane_ransac_ctx* c, const float* coeffs, const pr_point* border, const size_t* border_offsets,
                          int n_planes, float dist_threshold, unsigned rand_seed, int32_t* absorbed_cur, int32_t* absorbed_orig,
                          size_t idx_cap, size_t* plane_offsets, size_t* n_remaining) {
  PR_TRY(check_ctx(c));
  if (c->profiling) collect_spans(c);
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (n_planes < 0 || (n_planes && (!coeffs || !border || !border_offsets))) return fail(PR_ERR_INVALID, "bad plane arguments");
  if (!plane_offsets) return fail(PR_ERR_INVALID, "null plane_offsets");
  if (std::isnan(dist_threshold)) return fail(PR_ERR_INVALID, "dist_threshold is NaN");
  const size_t P = (size_t)n_planes;
  size_t n_border = 0;
  for (size_t j = 0; j < P; ++j) {
    if (border_offsets[j + 1] <= border_offsets[j]) return fail(PR_ERR_INVALID, "plane %zu has an empty border polygon", j);
    if (border_offsets[j + 1] - border_offsets[j] > (size_t)INT_MAX / 16) return fail(PR_ERR_INVALID, "border polygon %zu too large", j);
  }
  if (P) n_border = border_offsets[P] - border_offsets[0];
  if (n_border > (size_t)INT_MAX / 4) return fail(PR_ERR_INVALID, "too many border vertices");
  for (size_t j = 0; j <= P; ++j) plane_offsets[j] = 0;
  const size_t n = c->n_current;
  if (n_remaining) *n_remaining = n;
  if (P == 0 || n == 0) return PR_OK;
  PR_TRY(reserve_small(c));
  PR_TRY(reserve_work(c));

  // per-call constants: planes, border vertices, the ten edge indices rand() % border.size() draws after srand(seed)
  std::vector<pr::ReabsorbPlane> hp(P);
  std::vector<int32_t> hre(10 * P);
  for (size_t j = 0; j < P; ++j) {
    hp[j].a = coeffs[4 * j]; hp[j].b = coeffs[4 * j + 1]; hp[j].c = coeffs[4 * j + 2]; hp[j].d = coeffs[4 * j + 3];
    hp[j].border_begin = (int)(border_offsets[j] - border_offsets[0]);
    hp[j].border_size = (int)(border_offsets[j + 1] - border_offsets[j]);
    hp[j].pad0 = hp[j].pad1 = 0;
    pr::msvc_rand_edges(rand_seed, hp[j].border_size, &hre[10 * j]);
  }
  PR_TRY(dev_reserve(c->d_rb_planes, P));
  PR_TRY(dev_reserve(c->d_rb_ray_edges, 10 * P));
  PR_TRY(dev_reserve(c->d_rb_border, n_border));
  PR_TRY(dev_reserve(c->d_rb_edges, 3 * n_border));
  PR_TRY(dev_reserve(c->d_rb_rays, 10 * P));
  PR_TRY(dev_reserve(c->d_rb_counts, P));
  PR_TRY(dev_reserve(c->d_rb_counters, 2));
  PR_TRY(dev_reserve(c->d_rb_claimed, c->current.cap));
  PR_CUDA(cudaMemcpyAsync(c->d_rb_planes.p, hp.data(), P * sizeof(pr::ReabsorbPlane), cudaMemcpyHostToDevice, c->stream));
  PR_CUDA(cudaMemcpyAsync(c->d_rb_ray_edges.p, hre.data(), 10 * P * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  PR_CUDA(cudaMemcpyAsync(c->d_rb_border.p, border + border_offsets[0], n_border * sizeof(pr_point), cudaMemcpyHostToDevice, c->stream));
  PR_CUDA(cudaMemsetAsync(c->d_rb_counts.p, 0, P * sizeof(int32_t), c->stream));
  PR_CUDA(cudaMemsetAsync(c->d_rb_counters.p, 0, 2 * sizeof(unsigned long long), c->stream));
  PR_CUDA(cudaMemsetAsync(c->d_rb_claimed.p, 0, c->current.cap * sizeof(uint32_t), c->stream));
  {
    Span sp(c, KC_OTHER, 1);
    pr::launch_reabsorb_prepare(c->d_rb_border.p, c->d_rb_planes.p, n_planes, c->d_rb_ray_edges.p, c->d_rb_edges.p, c->d_rb_rays.p, c->stream);
  }
  // R1: candidate pairs (retry once with the exact size when the first guess was too small)
  unsigned long long h_counters[2] = {0, 0};
  size_t cand_cap = std::max<size_t>(c->d_rb_cand.cap, std::max<size_t>(n, (size_t)1 << 16));
  for (int attempt = 0; attempt < 2; ++attempt) {
    PR_TRY(dev_reserve(c->d_rb_cand, cand_cap));
    {
      Span sp(c, KC_OTHER, 1);
      pr::launch_reabsorb_filter(c->current, n, c->d_rb_planes.p, n_planes, dist_threshold, c->d_rb_counters.p, c->d_rb_cand.p,
                                 (unsigned long long)c->d_rb_cand.cap, c->num_sms, c->stream);
    }
    PR_CUDA(cudaGetLastError());
    PR_CUDA(cudaMemcpyAsync(h_counters, c->d_rb_counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, c->stream));
    PR_TRY(sync_stream(c));
    if (h_counters[0] <= (unsigned long long)c->d_rb_cand.cap) break;
    if (attempt == 1) return fail(PR_ERR_CUDA, "candidate count changed between passes");
    cand_cap = (size_t)h_counters[0];
    PR_CUDA(cudaMemsetAsync(c->d_rb_counters.p, 0, sizeof(unsigned long long), c->stream));
  }
  const unsigned long long n_cand = h_counters[0];
  std::vector<int32_t> h_counts(P, 0);
  unsigned long long n_abs = 0;
  if (n_cand) {
    // R2: polygon containment
    PR_TRY(dev_reserve(c->d_rb_keys, 2 * (size_t)n_cand));
    {
      Span sp(c, KC_OTHER, 1);
      pr::launch_reabsorb_poly(c->current, c->d_rb_cand.p, n_cand, c->d_rb_planes.p, c->d_rb_edges.p, c->d_rb_rays.p, c->d_rb_claimed.p,
                               c->d_rb_keys.p, c->d_rb_counters.p + 1, c->d_rb_counts.p, c->num_sms, c->stream);
    }
    PR_CUDA(cudaGetLastError());
    PR_CUDA(cudaMemcpyAsync(h_counters, c->d_rb_counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, c->stream));
    PR_CUDA(cudaMemcpyAsync(h_counts.data(), c->d_rb_counts.p, P * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    PR_TRY(sync_stream(c));
    n_abs = h_counters[1];
  }
  for (size_t j = 0; j < P; ++j) plane_offsets[j + 1] = plane_offsets[j] + (size_t)h_counts[j];
  if ((absorbed_cur || absorbed_orig) && n_abs > idx_cap)
    return fail(PR_ERR_CAPACITY, "absorbed index buffers hold %zu entries, need %llu", idx_cap, n_abs);
  if (n_abs) {
    // R3: per-plane ascending lists
    int key_bits = 33;
    while (key_bits < 64 && (P >> (key_bits - 32)) != 0) ++key_bits;
    const size_t tb = pr::reabsorb_sort_temp_bytes((size_t)n_abs);
    PR_TRY(dev_reserve(c->d_rb_temp, tb + 256));
    PR_TRY(dev_reserve(c->d_rb_out_cur, (size_t)n_abs));
    PR_TRY(dev_reserve(c->d_rb_out_orig, (size_t)n_abs));
    {
      Span sp(c, KC_OTHER, 3);
      pr::launch_reabsorb_lists(c->d_rb_keys.p, c->d_rb_keys.p + n_cand, (size_t)n_abs, key_bits, c->d_rb_temp.p, tb, c->current.orig,
                                c->d_rb_out_cur.p, c->d_rb_out_orig.p, c->stream);
    }
    PR_CUDA(cudaGetLastError());
    if (absorbed_cur) PR_CUDA(cudaMemcpyAsync(absorbed_cur, c->d_rb_out_cur.p, n_abs * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (absorbed_orig) PR_CUDA(cudaMemcpyAsync(absorbed_orig, c->d_rb_out_orig.p, n_abs * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  // peel: the unclaimed points become the current cloud (":1560-1566")
  pr::CloudView dst = (c->current.x == c->work[0].x) ? c->work[1] : c->work[0];
  PR_TRY(dev_reserve(c->d_scratch, pr::compact_scratch_bytes(n) + 64));
  {
    Span sp(c, KC_COMPACT, 1);
    pr::Plane4 none = {0, 0, 0, 0};
    pr::launch_compact(c->current, n, none, 0.f, 3, dst, true, nullptr, nullptr, c->d_scratch.p, c->d_totals.p, c->stream, c->d_rb_claimed.p);
    c->prof.points_compact += (long long)n;
  }
  PR_CUDA(cudaGetLastError());
  if (c->comm) PR_NCCL(g_nccl.AllGather(c->d_totals.p, c->d_totals.p + 2, 2, ncclInt64, c->comm, c->stream));
  PR_CUDA(cudaMemcpyAsync(c->h_totals.p, c->d_totals.p, (2 + (c->comm ? 2 * c->n_ranks : 0)) * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  PR_TRY(sync_stream(c));
  c->prof.bytes_compact += 4ll * (long long)n + compact_bytes(c->current, n, true, c->h_totals.p[0], 0, false, false);
  c->current = dst;
  c->n_current = (size_t)c->h_totals.p[0];
  if (c->comm) {
    long long tot = 0, f = 0;
    for (int r = 0; r < c->n_ranks; ++r) {
      if (r == c->rank) f = tot;
      tot += c->h_totals.p[2 + 2 * r];
    }
    c->n_global_current = tot;
    c->first_current = f;
  } else {
    c->n_global_current = (long long)c->n_current;
    c->first_current = 0;
  }
  if (n_remaining) *n_remaining = c->n_current;
  return PR_OK;
}

// ---- batch of small clouds --------------------------------------------------------------------
int plane_ransac_set_cloud_batch(plane_ransac_ctx* c, const pr_point* pts, size_t n_clouds, size_t n_per_cloud) {
  PR_TRY(check_ctx(c));
  if (n_clouds == 0 || n_per_cloud == 0 || !pts) return fail(PR_ERR_INVALID, "empty batch");
  if (n_per_cloud > (size_t)INT_MAX - 4096 || n_clouds > (size_t)1 << 24) return fail(PR_ERR_INVALID, "batch too large");
  const size_t stride = (n_per_cloud + pr::kTilePoints - 1) / pr::kTilePoints * pr::kTilePoints;
  const size_t cap = n_clouds * stride + pr::kTilePoints;
  const size_t n_total = n_clouds * n_per_cloud;
  PR_TRY(dev_reserve(c->aos, n_total));
  PR_TRY(dev_reserve(c->batch_mem, 3 * cap));
  PR_TRY(dev_reserve(c->d_batch_bbox, 6 * n_clouds));
  PR_CUDA(cudaMemcpyAsync(c->aos.p, pts, n_total * sizeof(pr_point), cudaMemcpyHostToDevice, c->stream));
  c->batch_view = planes_view(c->batch_mem.p, nullptr, cap);
  {
    Span sp(c, KC_STAGE, 2);
    pr::launch_stage_batch(c->aos.p, n_clouds, n_per_cloud, stride, c->batch_view, c->d_batch_bbox.p, c->stream);
  }
  PR_CUDA(cudaGetLastError());
  std::vector<uint32_t> keys(6 * n_clouds);
  PR_CUDA(cudaMemcpyAsync(keys.data(), c->d_batch_bbox.p, keys.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  PR_CUDA(cudaStreamSynchronize(c->stream));
  c->batch_scale_exp.resize(n_clouds);
  for (size_t i = 0; i < n_clouds; ++i) c->batch_scale_exp[i] = pr::scale_exp_from_bbox_keys(&keys[6 * i]);
  {
    std::vector<double> scales(n_clouds);
    for (size_t i = 0; i < n_clouds; ++i) scales[i] = std::ldexp(1.0, c->batch_scale_exp[i]);
    PR_TRY(dev_reserve(c->d_batch_sexp, n_clouds));
    PR_TRY(dev_reserve(c->d_batch_scale, n_clouds));
    PR_CUDA(cudaMemcpyAsync(c->d_batch_sexp.p, c->batch_scale_exp.data(), n_clouds * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PR_CUDA(cudaMemcpyAsync(c->d_batch_scale.p, scales.data(), n_clouds * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
  }
  c->batch_clouds = n_clouds;
  c->batch_n = n_per_cloud;
  c->batch_stride = stride;
  return PR_OK;
}

}  // extern "C"

namespace {

// Every cloud's ascending inlier list for d_planes (clouds with ok[c] < 0 have none), in two halves so that a caller can
// put its own read-backs between them under one synchronisation: queue the offsets + list kernels and the offsets'
// read-back; then, once the stream has been synchronised, check the capacity and fetch the lists.
int batch_lists_queue(plane_ransac_ctx* c, const pr_params* prm, const float4* d_planes, const int32_t* d_ok, const int32_t* d_cnt,
                      int32_t* inliers, size_t cap, size_t* offsets) {
  const size_t C = c->batch_clouds, n = c->batch_n;
  if (!offsets) return PR_OK;
  const float t = pr::threshold_up(prm->distance_threshold);
  PR_TRY(dev_reserve(c->d_batch_offs, C + 1));
  const size_t dev_cap = inliers ? std::min(cap, C * n) : 0;
  if (dev_cap) PR_TRY(dev_reserve(c->d_batch_lists, dev_cap));
  {
    Span sp(c, KC_COMPACT, dev_cap ? 2 : 1);
    pr::launch_batch_lists(c->batch_view, n, c->batch_stride, (int)C, d_planes, d_ok, t, prm->dot_order, d_cnt, c->d_batch_offs.p, dev_cap,
                           dev_cap ? c->d_batch_lists.p : nullptr, c->stream);
  }
  PR_CUDA(cudaGetLastError());
  static_assert(sizeof(size_t) == sizeof(unsigned long long), "size_t is 64 bits");
  PR_CUDA(cudaMemcpyAsync(offsets, c->d_batch_offs.p, (C + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  return PR_OK;
}

int batch_lists_fetch(plane_ransac_ctx* c, int32_t* inliers, size_t cap, const size_t* offsets) {
  if (!offsets) return PR_OK;
  const size_t total = offsets[c->batch_clouds];
  if (inliers && total > cap) return fail(PR_ERR_CAPACITY, "inlier buffer holds %zu entries, the batch has %zu inliers", cap, total);
  if (inliers && total) {
    PR_CUDA(cudaMemcpyAsync(inliers, c->d_batch_lists.p, total * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    PR_TRY(sync_stream(c));
  }
  return PR_OK;
}

// Score-all mode without the host in the loop: every step of segment() for all clouds is a launch; one read-back at the
// end.  *fell_back: a degenerate sample somewhere needs PCL's redraw rule — the host-driven path redoes the batch.
int segment_batch_device(plane_ransac_ctx* c, const pr_params* prm, float* coeffs, int32_t* n_inliers, int32_t* inliers, size_t cap,
                         size_t* offsets, pr_segment_info* infos, bool* fell_back) {
  HostTimer whole(&c->prof.host_ms_total);
  const size_t C = c->batch_clouds, n = c->batch_n, stride = c->batch_stride;
  const int K = prm->max_iterations + 1;
  const size_t CK = C * (size_t)K;
  const float t = pr::threshold_up(prm->distance_threshold);
  *fell_back = false;
  if (c->batch_tri.size() != 3 * (size_t)K || c->batch_tri_n != n || c->batch_tri_seed != prm->seed) {
    // every cloud has n points and the same seed, so PCL draws the same triples for every cloud: drawn once and kept
    HostTimer tm(&c->prof.host_ms_sampling);
    c->batch_tri.resize(3 * (size_t)K);
    pr::IndexSampler sampler(n, prm->seed);
    for (int j = 0; j < K; ++j) sampler.draw(c->batch_tri.data() + 3 * j);
    c->batch_tri_n = n;
    c->batch_tri_seed = prm->seed;
    c->batch_tri_dev_valid = 0;
  }
  PR_TRY(reserve_draws(c, CK, false));
  // every per-cloud result in ONE device block (a single read-back): flag | raw | refined | best | best count | final count
  const size_t words = 4 + 8 * C + 3 * C;  // 32-bit words; the float4 parts stay 16-byte aligned
  PR_TRY(dev_reserve(c->d_batch_out, words));
  PR_TRY(pin_reserve(c->h_batch_out, words));
  int* d_flag = reinterpret_cast<int*>(c->d_batch_out.p);
  float4* d_raw = reinterpret_cast<float4*>(c->d_batch_out.p + 4);
  float4* d_refined = d_raw + C;
  int32_t* d_best = reinterpret_cast<int32_t*>(d_refined + C);
  int32_t* d_bestcnt = d_best + C;
  int32_t* d_cnt = prm->optimize_coefficients ? d_bestcnt + C : d_bestcnt;  // without the refit the raw model's count is final
  PR_TRY(dev_reserve(c->d_batch_refit, C));
  if (c->batch_tri_dev_valid != 3 * (size_t)K) {  // the triples only change with (n, seed, K): uploaded once
    std::memcpy(c->h_triples.p, c->batch_tri.data(), 3 * (size_t)K * sizeof(int32_t));
    PR_TRY(dev_reserve(c->d_batch_tri, 3 * (size_t)K));
    PR_CUDA(cudaMemcpyAsync(c->d_batch_tri.p, c->h_triples.p, 3 * (size_t)K * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
    c->batch_tri_dev_valid = 3 * (size_t)K;
  }
  // One dependent chain of launches (each set up while its predecessor drains: pr_chain_dev.cuh), no memsets between
  // them: the models launch clears the counts and the flag, the decision launch clears the moments.
  {
    Span sp(c, KC_MODELS, 1);
    pr::launch_batch_gather_models(c->batch_view, n, stride, (int)C, c->d_batch_tri.p, K, c->d_sample_pts.p, c->d_hyps.p, c->d_good.p,
                                   c->d_counts.p, d_flag, c->stream);
  }
  {
    Span sp(c, KC_SCORE, 0);
    c->prof.launches_score += pr::launch_score(c->batch_view, n, (int)C, stride, c->d_hyps.p, K, t, prm->dot_order, c->d_counts.p, c->num_sms, c->stream,
                                               nullptr, true);
    c->prof.pairs_scored += (long long)n * (long long)CK;
  }
  {
    Span sp(c, KC_OTHER, 1);
    pr::launch_batch_replay(c->d_counts.p, c->d_good.p, K, (int)C, d_best, d_bestcnt, d_flag,
                            prm->optimize_coefficients ? c->d_batch_refit.p : nullptr, c->stream);
  }
  if (prm->optimize_coefficients) {
    Span sp(c, KC_REFIT, 1);
    pr::launch_refit_batch(c->batch_view, n, stride, (int)C, c->d_hyps.p, c->d_sample_pts.p, K, d_best, t, prm->dot_order,
                           c->d_batch_scale.p, c->d_batch_refit.p, c->stream, true);
    c->prof.points_refit += (long long)(n * C);
    c->prof.bytes_refit += 12ll * (long long)(n * C);
  }
  {
    Span sp(c, KC_OTHER, 1);
    pr::launch_batch_finish(c->d_hyps.p, K, d_best, c->d_batch_refit.p, c->d_batch_sexp.p, prm->optimize_coefficients ? 1 : 0,
                            (int)C, d_raw, d_refined, c->stream);
  }
  if (prm->optimize_coefficients) {
    Span sp(c, KC_COMPACT, 1);
    pr::launch_batch_count(c->batch_view, n, stride, (int)C, d_refined, d_best, t, prm->dot_order, d_cnt, c->stream);
  }
  PR_CUDA(cudaGetLastError());
  // the lists are queued before anything is read back (a batch that has to be redone wastes them, nothing else)
  PR_TRY(batch_lists_queue(c, prm, d_refined, d_best, d_cnt, inliers, cap, offsets));
  PR_CUDA(cudaMemcpyAsync(c->h_batch_out.p, c->d_batch_out.p, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  PR_TRY(sync_stream(c));
  const int flag = *reinterpret_cast<const int*>(c->h_batch_out.p);
  const float4* raw = reinterpret_cast<const float4*>(c->h_batch_out.p + 4);
  const float4* refined = raw + C;
  const int32_t* best = reinterpret_cast<const int32_t*>(refined + C);
  const int32_t* best_cnt = best + C;
  const int32_t* final_cnt = prm->optimize_coefficients ? best_cnt + C : best_cnt;
  if (flag) {
    *fell_back = true;
    return PR_OK;
  }
  for (size_t i = 0; i < C; ++i) {
    n_inliers[i] = final_cnt[i];
    const float r4[4] = {refined[i].x, refined[i].y, refined[i].z, refined[i].w};
    std::memcpy(coeffs + 4 * i, r4, sizeof(r4));
    if (infos) {
      pr_segment_info inf;
      std::memset(&inf, 0, sizeof(inf));
      inf.ok = 1;
      inf.iterations = inf.draws = inf.n_scored = K;
      inf.n_cloud = (long long)n;
      inf.scale_exp = c->batch_scale_exp[i];
      for (int k = 0; k < 3; ++k) inf.best_sample[k] = c->batch_tri[3 * (size_t)best[i] + k];
      inf.best_count = inf.n_inliers_raw = best_cnt[i];
      inf.raw_coeff[0] = raw[i].x; inf.raw_coeff[1] = raw[i].y; inf.raw_coeff[2] = raw[i].z; inf.raw_coeff[3] = raw[i].w;
      inf.n_inliers = final_cnt[i];
      infos[i] = inf;
    }
  }
  return batch_lists_fetch(c, inliers, cap, offsets);
}

int segment_batch_host(plane_ransac_ctx* c, const pr_params* prm, float* coeffs, int32_t* n_inliers, int32_t* inliers, size_t cap,
                       size_t* offsets, pr_segment_info* infos) {
  HostTimer whole(&c->prof.host_ms_total);
  const size_t C = c->batch_clouds, n = c->batch_n, stride = c->batch_stride;
  const float t = pr::threshold_up(prm->distance_threshold);
  std::vector<pr::RansacReplay> replay;
  replay.reserve(C);
  for (size_t i = 0; i < C; ++i) replay.emplace_back((long long)std::max<size_t>(n, 1), prm->max_iterations, prm->probability);
  // every cloud has n points and the same seed, so PCL draws the same triples for every cloud
  std::vector<int32_t> all_triples;
  int total_draws = 0;
  auto all_done = [&] {
    for (auto& r : replay)
      if (!r.done()) return false;
    return true;
  };
  if (n >= 3 && !all_done()) {
    pr::IndexSampler sampler(n, prm->seed);
    long long prev = 0;
    while (!all_done()) {
      int min_it = INT_MAX;
      for (auto& r : replay)
        if (!r.done()) min_it = std::min(min_it, r.iterations());
      const long long trials_left = (long long)prm->max_iterations + 1 - min_it;
      long long B = prm->probability >= 1.0 ? trials_left : prev == 0 ? std::min<long long>(trials_left, 256)
                                                                      : std::min<long long>(trials_left, 2 * prev);
      if (B < 1) B = 1;
      if ((double)B * (double)C > 4.0e8) B = std::max<long long>(1, (long long)(4.0e8 / (double)C));
      prev = B;
      const size_t CB = C * (size_t)B;
      PR_TRY(reserve_draws(c, CB, false));  // sample_pts 3*CB, hyps CB, counts CB, good CB (triples 3*CB, over-sized)
      all_triples.resize(3 * ((size_t)total_draws + (size_t)B));
      int32_t* ht = all_triples.data() + 3 * (size_t)total_draws;
      {
        HostTimer tm(&c->prof.host_ms_sampling);
        for (long long j = 0; j < B; ++j) sampler.draw(ht + 3 * j);
      }
      std::memcpy(c->h_triples.p, ht, 3 * (size_t)B * sizeof(int32_t));
      PR_CUDA(cudaMemcpyAsync(c->d_triples.p, c->h_triples.p, 3 * (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
      {
        Span sp(c, KC_MODELS, 2);
        pr::launch_gather_samples(c->batch_view, 0, n, c->d_triples.p, (int)(3 * B), c->d_sample_pts.p, (int)C, stride, c->stream);
        pr::launch_models(c->d_sample_pts.p, (int)CB, c->d_hyps.p, c->d_good.p, c->stream);
      }
      PR_CUDA(cudaMemsetAsync(c->d_counts.p, 0, CB * sizeof(int32_t), c->stream));
      {
        Span sp(c, KC_SCORE, 0);
        c->prof.launches_score += pr::launch_score(c->batch_view, n, (int)C, stride, c->d_hyps.p, (int)B, t, prm->dot_order, c->d_counts.p, c->num_sms, c->stream);
        c->prof.pairs_scored += (long long)n * (long long)CB;
      }
      PR_CUDA(cudaGetLastError());
      PR_CUDA(cudaMemcpyAsync(c->h_counts.p, c->d_counts.p, CB * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
      PR_CUDA(cudaMemcpyAsync(c->h_good.p, c->d_good.p, CB * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
      PR_TRY(sync_stream(c));
      {
        HostTimer tm(&c->prof.host_ms_replay);
        std::vector<uint8_t> good8((size_t)B);
        for (size_t i = 0; i < C; ++i) {
          if (replay[i].done()) continue;
          const int32_t* g = c->h_good.p + i * (size_t)B;
          for (long long j = 0; j < B; ++j) good8[j] = g[j] ? 1 : 0;
          replay[i].feed(c->h_counts.p + i * (size_t)B, good8.data(), (int)B);
        }
      }
      total_draws += (int)B;
    }
  }

  // winners: per-cloud best triple -> model + pivot again (K = 1 per cloud), refit, final count
  PR_TRY(dev_reserve(c->d_batch_idx, 3 * C));
  PR_TRY(dev_reserve(c->d_batch_pts, 3 * C));
  PR_TRY(dev_reserve(c->d_batch_hyps, C));
  PR_TRY(dev_reserve(c->d_batch_cnt, 2 * C));
  PR_TRY(dev_reserve(c->d_batch_scale, C));
  PR_TRY(dev_reserve(c->d_batch_refit, C));
  std::vector<int32_t> best_tri(3 * C, 0), model_idx(C, -1);
  std::vector<double> scales(C);
  for (size_t i = 0; i < C; ++i) {
    const int b = replay[i].best_draw();
    if (b >= 0) {
      model_idx[i] = 0;
      for (int k = 0; k < 3; ++k) best_tri[3 * i + k] = all_triples[3 * (size_t)b + k];
    }
    scales[i] = std::ldexp(1.0, c->batch_scale_exp[i]);
  }
  std::vector<float4> raw(C);
  std::vector<pr::RefitOut> mom(C);
  PR_CUDA(cudaMemcpyAsync(c->d_batch_idx.p, best_tri.data(), 3 * C * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  {
    Span sp(c, KC_MODELS, 2);
    pr::launch_gather_samples(c->batch_view, 0, n, c->d_batch_idx.p, 3, c->d_batch_pts.p, (int)C, stride, c->stream, true);
    pr::launch_models(c->d_batch_pts.p, (int)C, c->d_batch_hyps.p, c->d_batch_cnt.p + C, c->stream);
  }
  PR_CUDA(cudaMemcpyAsync(raw.data(), c->d_batch_hyps.p, C * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  if (prm->optimize_coefficients) {
    PR_CUDA(cudaStreamSynchronize(c->stream));  // best_tri must be consumed before model_idx reuses the buffer
    PR_CUDA(cudaMemcpyAsync(c->d_batch_idx.p, model_idx.data(), C * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    PR_CUDA(cudaMemcpyAsync(c->d_batch_scale.p, scales.data(), C * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PR_CUDA(cudaMemsetAsync(c->d_batch_refit.p, 0, C * sizeof(pr::RefitOut), c->stream));
    {
      Span sp(c, KC_REFIT, 1);
      pr::launch_refit_batch(c->batch_view, n, stride, (int)C, c->d_batch_hyps.p, c->d_batch_pts.p, 1, c->d_batch_idx.p, t,
                             prm->dot_order, c->d_batch_scale.p, c->d_batch_refit.p, c->stream);
      c->prof.points_refit += (long long)(n * C);
      c->prof.bytes_refit += 12ll * (long long)(n * C);
    }
    PR_CUDA(cudaMemcpyAsync(mom.data(), c->d_batch_refit.p, C * sizeof(pr::RefitOut), cudaMemcpyDeviceToHost, c->stream));
  }
  PR_CUDA(cudaGetLastError());
  PR_TRY(sync_stream(c));
  std::vector<float4> refined(C);
  {
    HostTimer tm(&c->prof.host_ms_replay);
    for (size_t i = 0; i < C; ++i) {
      refined[i] = raw[i];
      if (prm->optimize_coefficients && model_idx[i] >= 0) {
        int64_t m[16];
        for (int k = 0; k < 16; ++k) m[k] = (int64_t)mom[i].m[k];
        float out[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
        pr::plane_from_moments(m, mom[i].pivot, c->batch_scale_exp[i], out);
        refined[i] = make_float4(out[0], out[1], out[2], out[3]);
      }
    }
  }
  std::vector<int32_t> final_cnt(C, 0);
  if (prm->optimize_coefficients) {
    PR_CUDA(cudaMemcpyAsync(c->d_batch_hyps.p, refined.data(), C * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    PR_CUDA(cudaMemsetAsync(c->d_batch_cnt.p, 0, C * sizeof(int32_t), c->stream));
    {
      Span sp(c, KC_SCORE, 0);
      c->prof.launches_score += pr::launch_score(c->batch_view, n, (int)C, stride, c->d_batch_hyps.p, 1, t, prm->dot_order, c->d_batch_cnt.p, c->num_sms, c->stream);
      c->prof.pairs_scored += (long long)(n * C);
    }
    PR_CUDA(cudaGetLastError());
    PR_CUDA(cudaMemcpyAsync(final_cnt.data(), c->d_batch_cnt.p, C * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    PR_TRY(sync_stream(c));
  }
  for (size_t i = 0; i < C; ++i) {
    const bool ok = model_idx[i] >= 0;
    const int cnt = !ok ? 0 : prm->optimize_coefficients ? final_cnt[i] : replay[i].best_count();
    n_inliers[i] = cnt;
    const float z4[4] = {0, 0, 0, 0};
    const float r4[4] = {refined[i].x, refined[i].y, refined[i].z, refined[i].w};
    std::memcpy(coeffs + 4 * i, ok ? r4 : z4, 4 * sizeof(float));
    if (infos) {
      pr_segment_info inf;
      std::memset(&inf, 0, sizeof(inf));
      inf.ok = ok ? 1 : 0;
      inf.iterations = replay[i].iterations();
      inf.draws = replay[i].draws_used();
      inf.skipped = replay[i].skipped();
      inf.n_scored = total_draws;
      inf.n_cloud = (long long)n;
      inf.scale_exp = c->batch_scale_exp[i];
      if (ok) {
        for (int k = 0; k < 3; ++k) inf.best_sample[k] = best_tri[3 * i + k];
        inf.best_count = replay[i].best_count();
        inf.n_inliers_raw = inf.best_count;
        inf.raw_coeff[0] = raw[i].x; inf.raw_coeff[1] = raw[i].y; inf.raw_coeff[2] = raw[i].z; inf.raw_coeff[3] = raw[i].w;
        inf.n_inliers = cnt;
      }
      infos[i] = inf;
    }
  }
  if (!offsets) return PR_OK;
  // index lists: the final planes and a per-cloud "has a model" flag on the device, then the same list kernels
  PR_CUDA(cudaMemcpyAsync(c->d_batch_hyps.p, refined.data(), C * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  PR_CUDA(cudaMemcpyAsync(c->d_batch_idx.p, model_idx.data(), C * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  PR_CUDA(cudaMemcpyAsync(c->d_batch_cnt.p, n_inliers, C * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  PR_TRY(batch_lists_queue(c, prm, c->d_batch_hyps.p, c->d_batch_idx.p, c->d_batch_cnt.p, inliers, cap, offsets));
  PR_TRY(sync_stream(c));
  return batch_lists_fetch(c, inliers, cap, offsets);
}

}  // namespace

extern "C" {

int plane_ransac_segment_batch_lists(plane_ransac_ctx* c, const pr_params* prm, float* coeffs, int32_t* n_inliers, int32_t* inliers,
                                     size_t cap, size_t* offsets, pr_segment_info* infos) {
  PR_TRY(check_ctx(c));
  if (c->profiling) collect_spans(c);
  PR_TRY(check_params(prm));
  if (c->batch_clouds == 0) return fail(PR_ERR_NO_CLOUD, "no batch staged");
  if (!coeffs || !n_inliers) return fail(PR_ERR_INVALID, "null output");
  if (inliers && !offsets) return fail(PR_ERR_INVALID, "inlier lists need the offsets array");
  if (prm->optimize_coefficients && prm->refit_mode != PR_REFIT_FIXED)
    return fail(PR_ERR_INVALID, "PR_REFIT_PCL_FLOAT is not available for batches (use plane_ransac_segment_one per cloud)");
  const bool device_loop = c->round_loop != PR_LOOP_HOST && prm->probability >= 1.0 && prm->max_iterations >= 1 && c->batch_n >= 3 &&
                           (double)c->batch_clouds * ((double)prm->max_iterations + 1.0) <= 4.0e8;
  if (device_loop) {
    bool fell_back = false;
    PR_TRY(segment_batch_device(c, prm, coeffs, n_inliers, inliers, cap, offsets, infos, &fell_back));
    if (!fell_back) return PR_OK;
  }
  return segment_batch_host(c, prm, coeffs, n_inliers, inliers, cap, offsets, infos);
}

int plane_ransac_segment_batch(plane_ransac_ctx* c, const pr_params* prm, float* coeffs, int32_t* n_inliers, pr_segment_info* infos) {
  return plane_ransac_segment_batch_lists(c, prm, coeffs, n_inliers, nullptr, 0, nullptr, infos);
}

// ---- sharding -----------------------------------------------------------------------------------
int plane_ransac_comm_unique_id(void* out128) {
  if (!out128) return fail(PR_ERR_INVALID, "null output");
  PR_TRY(load_nccl());
  ncclUniqueId id;
  PR_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == PLANE_RANSAC_UNIQUE_ID_BYTES, "ncclUniqueId size");
  std::memcpy(out128, &id, sizeof(id));
  return PR_OK;
}

int plane_ransac_comm_init(plane_ransac_ctx* c, int n_ranks, int rank, const void* unique_id128) {
  PR_TRY(check_ctx(c));
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !unique_id128) return fail(PR_ERR_INVALID, "bad rank/n_ranks/id");
  if (c->comm) return fail(PR_ERR_INVALID, "communicator already initialised");
  PR_TRY(load_nccl());
  ncclUniqueId id;
  std::memcpy(&id, unique_id128, sizeof(id));
  PR_NCCL(g_nccl.CommInitRank(&c->comm, n_ranks, id, rank));
  c->n_ranks = n_ranks;
  c->rank = rank;
  PR_TRY(reserve_small(c));
  PR_TRY(p2p_setup(c));
  if (c->have_cloud) {
    PR_TRY(refresh_global(c));
    c->n_global_current = c->n_global_staged;
    c->first_current = c->first_staged;
  }
  return PR_OK;
}

int plane_ransac_comm_p2p_enabled(plane_ransac_ctx* c) { return c && c->p2p_on ? 1 : 0; }

int plane_ransac_shard_info(plane_ransac_ctx* c, long long* n_global_staged, long long* first_staged,
                            long long* n_global_current, long long* first_current) {
  if (!c) return fail(PR_ERR_INVALID, "null context");
  if (!c->have_cloud) return fail(PR_ERR_NO_CLOUD, "no cloud staged");
  if (n_global_staged) *n_global_staged = c->n_global_staged;
  if (first_staged) *first_staged = c->first_staged;
  if (n_global_current) *n_global_current = c->n_global_current;
  if (first_current) *first_current = c->first_current;
  return PR_OK;
}

int plane_ransac_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(PR_ERR_INVALID, "null output");
  *out = nullptr;
  if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return fail(PR_ERR_OOM, "cudaMallocHost of %zu bytes failed", bytes);
  }
  return PR_OK;
}

int plane_ransac_host_free(void* p) {
  if (p && pr::release_plain_alloc(p)) return PR_OK;  // a reader's buffer on a host without a CUDA device
  if (p && cudaFreeHost(p) != cudaSuccess) {
    cudaGetLastError();
    return fail(PR_ERR_CUDA, "cudaFreeHost failed");
  }
  return PR_OK;
}

int plane_ransac_host_register(void* p, size_t bytes) {
  if (!p || !bytes) return fail(PR_ERR_INVALID, "null or empty range");
  if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
    const cudaError_t e = cudaGetLastError();
    return fail(PR_ERR_CUDA, "cudaHostRegister of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  }
  return PR_OK;
}

int plane_ransac_host_unregister(void* p) {
  if (p && cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return fail(PR_ERR_CUDA, "cudaHostUnregister failed");
  }
  return PR_OK;
}

// ---- measurement ----------------------------------------------------------------------------------
int plane_ransac_profile_enable(plane_ransac_ctx* c, int on) {
  PR_TRY(check_ctx(c));
  collect_spans(c);
  c->profiling = on != 0;
  return PR_OK;
}

int plane_ransac_profile_reset(plane_ransac_ctx* c) {
  PR_TRY(check_ctx(c));
  collect_spans(c);
  std::memset(&c->prof, 0, sizeof(c->prof));
  if (c->d_p2p_wait.p) PR_CUDA(cudaMemsetAsync(c->d_p2p_wait.p, 0, 8 * sizeof(unsigned long long), c->stream));
  return PR_OK;
}

int plane_ransac_profile_get(plane_ransac_ctx* c, pr_profile* out) {
  PR_TRY(check_ctx(c));
  if (!out) return fail(PR_ERR_INVALID, "null output");
  collect_spans(c);
  *out = c->prof;
  if (c->d_p2p_wait.p) {
    unsigned long long w[8];
    PR_CUDA(cudaMemcpyAsync(w, c->d_p2p_wait.p, sizeof(w), cudaMemcpyDeviceToHost, c->stream));
    PR_CUDA(cudaStreamSynchronize(c->stream));
    for (int ch = 0; ch < 4; ++ch) {
      out->p2p_wait_ms[ch] = (double)w[2 * ch] * 1e-6;
      out->p2p_exchanges[ch] = (long long)w[2 * ch + 1];
    }
  }
  return PR_OK;
}

int plane_ransac_round_timeline(plane_ransac_ctx* c, unsigned long long* stamps, size_t cap_rounds, size_t* n_rounds) {
  PR_TRY(check_ctx(c));
  const size_t rounds = c->timeline.size() / pr::kStampSlots;
  if (n_rounds) *n_rounds = rounds;
  if (stamps) std::memcpy(stamps, c->timeline.data(), std::min(rounds, cap_rounds) * pr::kStampSlots * sizeof(unsigned long long));
  return PR_OK;
}

int plane_ransac_timer_start(plane_ransac_ctx* c) {
  PR_TRY(check_ctx(c));
  if (!c->timer_a) {
    PR_CUDA(cudaEventCreate(&c->timer_a));
    PR_CUDA(cudaEventCreate(&c->timer_b));
  }
  PR_CUDA(cudaEventRecord(c->timer_a, c->stream));
  return PR_OK;
}

int plane_ransac_timer_stop(plane_ransac_ctx* c, double* ms) {
  PR_TRY(check_ctx(c));
  if (!ms || !c->timer_a) return fail(PR_ERR_INVALID, "timer_stop without timer_start");
  PR_CUDA(cudaEventRecord(c->timer_b, c->stream));
  PR_CUDA(cudaEventSynchronize(c->timer_b));
  float t = 0.f;
  PR_CUDA(cudaEventElapsedTime(&t, c->timer_a, c->timer_b));
  *ms = t;
  return PR_OK;
}

int plane_ransac_measure_ffma_peak(plane_ransac_ctx* c, double* tflops) {
  PR_TRY(check_ctx(c));
  if (!tflops) return fail(PR_ERR_INVALID, "null output");
  const int grid = c->num_sms * 8, iters = 4096;
  DevBuf<float> out;
  PR_TRY(dev_reserve(out, (size_t)grid * 256));
  cudaEvent_t a, b;
  PR_CUDA(cudaEventCreate(&a));
  PR_CUDA(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    PR_CUDA(cudaEventRecord(a, c->stream));
    pr::launch_ffma_peak(out.p, iters, grid, c->stream);
    PR_CUDA(cudaEventRecord(b, c->stream));
    PR_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    PR_CUDA(cudaEventElapsedTime(&ms, a, b));
    // 8 chains x 8 unrolled FFMA2 per iteration, 2 FMA each, 2 FLOP per FMA
    const double flop = (double)grid * 256.0 * iters * 64.0 * 4.0;
    if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
  }
  c->prof.launches_other += 4;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  dev_free(out);
  *tflops = best;
  return PR_OK;
}

int plane_ransac_measure_copy_bw(plane_ransac_ctx* c, size_t bytes, double* gbs) {
  PR_TRY(check_ctx(c));
  if (!gbs) return fail(PR_ERR_INVALID, "null output");
  const size_t nvec = std::max<size_t>(bytes / 16, 1 << 20);
  DevBuf<float4> src, dst;
  PR_TRY(dev_reserve(src, nvec));
  PR_TRY(dev_reserve(dst, nvec));
  pr::launch_fill(src.p, nvec, 1.0f, c->num_sms, c->stream);
  cudaEvent_t a, b;
  PR_CUDA(cudaEventCreate(&a));
  PR_CUDA(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    PR_CUDA(cudaEventRecord(a, c->stream));
    pr::launch_copy(src.p, dst.p, nvec, c->num_sms, c->stream);
    PR_CUDA(cudaEventRecord(b, c->stream));
    PR_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    PR_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0) best = std::max(best, 2.0 * nvec * 16.0 / (ms * 1e-3) / 1e9);
  }
  c->prof.launches_other += 7;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  dev_free(src);
  dev_free(dst);
  *gbs = best;
  return PR_OK;
}

int plane_ransac_flush_l2(plane_ransac_ctx* c) {
  PR_TRY(check_ctx(c));
  const size_t nvec = (256u << 20) / 16;  // 256 MiB > 126 MB L2
  PR_TRY(dev_reserve(c->d_flush, nvec));
  pr::launch_fill(c->d_flush.p, nvec, 0.0f, c->num_sms, c->stream);
  c->prof.launches_other += 1;
  PR_CUDA(cudaStreamSynchronize(c->stream));
  return PR_OK;
}

// ---- host-side logic ------------------------------------------------------------------------------
int plane_ransac_host_draw_triples(size_t n_points, unsigned seed, int n_draws, int32_t* triples) {
  if (n_points < 3) return fail(PR_ERR_INVALID, "need at least 3 points to sample");
  if (n_draws < 0 || (n_draws && !triples)) return fail(PR_ERR_INVALID, "bad n_draws/triples");
  pr::IndexSampler s(n_points, seed);
  s.reserve((size_t)n_draws);
  for (int k = 0; k < n_draws; ++k) s.draw(triples + 3 * (size_t)k);
  return PR_OK;
}

int plane_ransac_host_draw_triples_parallel(size_t n_points, unsigned seed, int n_draws, int32_t* triples, int* fell_back) {
  if (n_points < 3 || n_points > (size_t)INT_MAX) return fail(PR_ERR_INVALID, "need 3 .. INT_MAX points to sample");
  if (n_draws < 0 || (n_draws && !triples) || !fell_back) return fail(PR_ERR_INVALID, "bad n_draws/triples");
  *fell_back = pr::draw_triples_parallel(n_points, seed, n_draws, triples) ? 0 : 1;
  return PR_OK;
}

int plane_ransac_host_replay(const int32_t* counts, const uint8_t* good, int n_draws, long long n_points,
                             int max_iterations, double probability, int* best_draw, int* iterations, int* draws_used,
                             int* skipped, int* exhausted) {
  if (n_draws < 0 || (n_draws && (!counts || !good)) || n_points < 1 || max_iterations < 0)
    return fail(PR_ERR_INVALID, "bad replay arguments");
  pr::RansacReplay r(n_points, max_iterations, probability);
  const bool done = r.done() || r.feed(counts, good, n_draws);
  if (best_draw) *best_draw = r.best_draw();
  if (iterations) *iterations = r.iterations();
  if (draws_used) *draws_used = r.draws_used();
  if (skipped) *skipped = r.skipped();
  if (exhausted) *exhausted = done ? 0 : 1;
  return PR_OK;
}

int plane_ransac_host_rand_edges(unsigned seed, int border_size, int32_t edges[10]) {
  if (border_size <= 0 || !edges) return fail(PR_ERR_INVALID, "border_size must be > 0");
  pr::msvc_rand_edges(seed, border_size, edges);
  return PR_OK;
}

int plane_ransac_host_shard_range(long long n_points, int n_ranks, int rank, long long* first, long long* count) {
  if (n_points < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !first || !count)
    return fail(PR_ERR_INVALID, "bad shard arguments");
  pr::shard_range(n_points, n_ranks, rank, first, count);
  return PR_OK;
}

int plane_ransac_host_plane_from_pcl_float_sums(const float sums[9], long long n_inliers, float coeff[4]) {
  if (!sums || !coeff) return fail(PR_ERR_INVALID, "null argument");
  return pr::plane_from_pcl_float_sums(sums, n_inliers, coeff) ? PR_OK : fail(PR_ERR_INVALID, "fewer than 4 inliers");
}

int plane_ransac_host_plane_from_moments(const int64_t m[16], const float pivot[3], int scale_exp, float coeff[4]) {
  if (!m || !pivot || !coeff) return fail(PR_ERR_INVALID, "null argument");
  return pr::plane_from_moments(m, pivot, scale_exp, coeff) ? PR_OK : fail(PR_ERR_INVALID, "fewer than 4 points in the moments");
}

}  // extern "C"
