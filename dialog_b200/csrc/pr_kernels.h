// pr_kernels.h — launch wrappers of the sm_100a kernels behind plane_ransac.h.
// Host-callable C++ (no CUDA syntax) so the orchestration in pr_api.cpp compiles with plain g++.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace pr {

// Points per scoring tile; every cloud plane is padded with NaN to a multiple of this, plus one
// spare tile, so tile loads never need a bounds check (NaN is never an inlier).
constexpr int kTilePoints = 1024;
// Points per compaction tile (one CTA).
constexpr int kCompactTile = 2048;

inline size_t padded_capacity(size_t n) {
  return (n + kTilePoints - 1) / kTilePoints * kTilePoints + kTilePoints;
}

// A cloud in HBM: three coordinate planes (+ optional original-index plane), each `cap` elements.
struct CloudView {
  float* x = nullptr;
  float* y = nullptr;
  float* z = nullptr;
  int32_t* orig = nullptr;  // nullptr == identity (the staged cloud)
  size_t cap = 0;
};

struct Plane4 { float a, b, c, d; };

// moments[16]: n, Sx, Sy, Sz, then (hi, lo) of Sxx, Sxy, Sxz, Syy, Syz, Szz; pivot in moments_pivot[3].
struct RefitOut {
  long long m[16];
  float pivot[4];
};

// ---- device-resident state of the peel loop (score-all mode): every kernel of a round reads its sizes, the winning
// draw and the refined plane from here, so a whole extraction is queued without the host in the loop (pr_chain.cu).
// Every kernel of a queued round notes %globaltimer once its predecessor has completed (thread 0 of block 0, right after
// griddepcontrol.wait); the stop rule copies the stamps into the round's record with the time it ran itself.  The
// difference of two consecutive stamps is a kernel plus the hand-over to the next one; no events, nothing on the host.
enum {
  kStampDraw = 0,   // sampler, parallel phase (also clears the round's accumulators)
  kStampResolve,    // sampler, collision replay
  kStampModels,     // sample points (+ their exchange) and computeModelCoefficients
  kStampScore,      // K2
  kStampDecide,     // (count exchange +) computeModel's decision
  kStampRefit,      // K3
  kStampFinish,     // closed-form plane as its own step: moment exchange, or rounds without the refit pass
  kStampPeel,       // K5
  kStampAdvance,    // remaining-count exchange + stop rule as its own step (sharded)
  kStampEnd,        // the round's record was written
  kStampSlots
};
struct RoundState {
  long long n_local;    // points of this rank's current cloud
  long long n_global;   // ... of all ranks
  long long first;      // global index of this rank's first point
  long long inl_off;    // inlier-list entries written so far (this rank)
  int round;            // planes accepted so far
  int stop;             // 0: running; 1: finished (plane too small / no model); 2: the round in flight goes back to the host loop
  int best;             // winning draw of the round in flight, -1: none
  int best_count;       // its inlier count
  float plane[4];       // coefficients of the round's final selection
  int pad[4];
  unsigned long long t[kStampSlots];  // %globaltimer when each kernel of the round in flight started (kStamp*)
};
// What one round decided; the host reads record r once round r has run (pr_segment_info + the peel bookkeeping).
struct RoundRecord {
  int ran;              // 0: the round was skipped (the loop had stopped before it)
  int accepted;         // the plane passed the minimum-size rule and was peeled
  int stop;             // RoundState::stop after the round
  int ok;               // a model was found
  int best, best_count;
  int best_sample[3];
  int n_draws;
  float raw[4], refined[4];
  long long n_cloud;                      // global size of the round's cloud
  long long n_local;                      // this rank's share of it
  long long n_inl_local, n_rem_local;     // this rank: inliers peeled / points left
  long long n_inl_global, n_rem_global;
  long long first_after;                  // this rank's first global index in the remaining cloud
  long long inl_off;                      // where this round's inlier lists start (this rank)
  unsigned long long t[kStampSlots];      // RoundState::t of the round, t[kStampEnd]: when the record was written
};

// On one GPU the last block of K3 / K5 also runs the step that follows it in the host-free loop (closed-form plane /
// stop rule), see pr_chain_dev.cuh; rec == nullptr: plain kernel.
struct ChainTail {
  RoundRecord* rec = nullptr;
  unsigned* ticket = nullptr;         // K3: blocks-done counter, zeroed by the round's prep kernel
  const int32_t* triples = nullptr;   // K3: the round's index triples (for the record)
  int scale_exp = 0, n_draws = 0;     // K3
  int min_plane = 0;                  // K5
};

// K0: AoS pr_point[n] -> planes (NaN padded to cap) + bounding box of the finite points.
// bbox: 6 x uint32 ordered-float encodings {min x,y,z, max x,y,z}; must be initialised by bbox_init.
void launch_bbox_init(uint32_t* bbox, cudaStream_t s);
void launch_stage(const float4* aos, size_t n, CloudView dst, uint32_t* bbox, cudaStream_t s);
// Batch of n_clouds equal-sized clouds: cloud c at element offset c * stride of each plane (stride = n_per
// rounded up to a tile), NaN between clouds; bbox: 6 keys per cloud.
void launch_stage_batch(const float4* aos, size_t n_clouds, size_t n_per, size_t stride, CloudView dst, uint32_t* bbox,
                        cudaStream_t s);
// Staging extras (reference preProcess): exact integer coordinate sums about the bbox corner `lo` on the
// 2^-scale_exp grid (sums[3] = number of finite points); in-place float translation by the centroid.
// launch_compact with dot_order == 2 removes the non-finite points (pcl::removeNaNFromPointCloud).
void launch_centroid(CloudView c, size_t n, const double lo[3], int scale_exp, long long* sums, int num_sms, cudaStream_t s);
void launch_translate(CloudView c, size_t n, const float centroid[3], int num_sms, cudaStream_t s);
// out[i] = cloud[idx[i]] (w = 1), optionally projected onto `pl` (reference projPoint2Plane arithmetic).
void launch_plane_points(CloudView cloud, const int32_t* idx, size_t n, Plane4 pl, bool project, float4* out, cudaStream_t s);
// out[i] = map ? map[idx[i]] : idx[i]
void launch_compose_map(const int32_t* idx, const int32_t* map, size_t n, int32_t* out, cudaStream_t s);
// planes -> AoS (w = 1)
void launch_unstage(CloudView src, size_t n, float4* aos, cudaStream_t s);

// K1a: sample_pts[s] = bits of point triples[s] if this rank owns it (first <= idx < first + n), else 0.
// clouds > 1: batch mode, the same triples are gathered from every cloud (cloud_stride elements apart), or
// with per_cloud_triples cloud c reads triples[c * n_samples ...].
void launch_gather_samples(CloudView cloud, long long first, size_t n, const int32_t* triples, int n_samples,
                           int4* sample_pts, int n_clouds, size_t cloud_stride, cudaStream_t s,
                           bool per_cloud_triples = false, const RoundState* st = nullptr);
// K1a + K1b in one launch (one GPU, host-free loop): gathers the three sample points of every triple from the cloud of
// st->n_local points into sample_pts and forms the models.
void launch_gather_models(CloudView cloud, const int32_t* triples, int n_models, int4* sample_pts, float4* hyps, int32_t* good,
                          const RoundState* st, cudaStream_t s);
// K1b: plane through each sample triple, PCL op order, no contraction; NaN plane + good = 0 when degenerate.
void launch_models(const int4* sample_pts, int n_models_total, float4* hyps, int32_t* good, cudaStream_t s);

// K2: counts[c * K + k] += |{ i in cloud c : |hyp[c*K+k] . (p_i, 1)| < t }|.  counts must be zeroed.
// Returns the number of kernel launches it made (K is cut into launches that fill their lane slots).
// st (single cloud only): the cloud size is st->n_local, read on the device (n_per_cloud is then only an upper bound), and
// the launch does nothing once st->stop is set.
int launch_score(CloudView cloud, size_t n_per_cloud, int n_clouds, size_t cloud_stride, const float4* hyps,
                 int K, float t, int dot_order, int32_t* counts, int num_sms, cudaStream_t s, const RoundState* st = nullptr,
                 bool chained = false);

// Hierarchical scorer: Morton-sorted copy of a cloud, per-32-point-block boxes, culled scoring with counts
// identical to launch_score.  keys / vals: 2*n uint32 each; temp: sort_temp_bytes(n); bounds: 2 float4 per block.
size_t sort_temp_bytes(size_t n);
void launch_morton_sort(CloudView src, size_t n, const float lo[3], float extent, uint32_t* keys, uint32_t* vals, void* temp,
                        size_t temp_bytes, CloudView dst, cudaStream_t s);
void launch_block_bounds(CloudView sorted, size_t n, float4* bounds, cudaStream_t s);
int launch_score_hier(CloudView sorted, size_t n, const float4* bounds, const float4* hyps, float2* aux, int K, float t,
                      float cmax, int dot_order, int32_t* counts, int num_sms, cudaStream_t s);

// K3: inlier predicate with hyps[model_index] + exact integer moments about the model's first sample point.
// out must be zeroed.  sample_pts supplies the pivot (sample_pts[3 * model_index]).
// st: n = st->n_local and model_index = st->best are read on the device (n is then an upper bound for the grid).
void launch_refit(CloudView cloud, size_t n, const float4* hyps, const int4* sample_pts, int model_index, float t,
                  int dot_order, int scale_exp, RefitOut* out, int num_sms, cudaStream_t s, RoundState* st = nullptr,
                  const ChainTail* tail = nullptr);

// PR_REFIT_PCL_FLOAT: out9 = PCL computeMeanAndCovarianceMatrix's nine FP32 sums (xx, xy, xz, yy, yz, zz, x, y, z) over the
// points idx[0 .. *n_idx_dev) of the cloud, added sequentially in that order by one thread.
void launch_refit_pcl_float(CloudView cloud, const int32_t* idx, const long long* n_idx_dev, float* out9, cudaStream_t s);

// K3 for a batch: cloud c refits hypothesis c * K + model_idx[c] (skipped when negative) with scale 2^s_c.
void launch_refit_batch(CloudView clouds, size_t n_per, size_t cloud_stride, int n_clouds, const float4* hyps,
                        const int4* sample_pts, int K, const int32_t* model_idx, float t, int dot_order,
                        const double* scales, RefitOut* outs, cudaStream_t s, bool chained = false);

// K5: stable partition by the inlier predicate of `plane`: remaining points -> dst (NaN re-padded),
// inlier positions -> inl_cur, their original indices -> inl_orig.  totals[0] = remaining, totals[1] = inliers.
// scratch: tile_state (n_tiles x uint64) + ticket, zeroed by the wrapper.
size_t compact_scratch_bytes(size_t n);
// dot_order == 2: predicate = point is non-finite (staging filter); dot_order == 3: predicate = flags[i] != 0
// (flags: one uint32 per point, as long as the cloud's capacity).
// st: n = st->n_local, the plane = st->plane and the list offset st->inl_off are read on the device (n is then an upper
// bound for the grid and the scratch size).
void launch_compact(CloudView src, size_t n, Plane4 plane, float t, int dot_order, CloudView dst, bool write_remaining,
                    int32_t* inl_cur, int32_t* inl_orig, void* scratch, long long* totals, cudaStream_t s,
                    const uint32_t* flags = nullptr, RoundState* st = nullptr, const ChainTail* tail = nullptr);

// Re-absorption pass of the reference's postProcessPlanes (pr_reabsorb.cu).
struct ReabsorbPlane {
  float a, b, c, d;
  int border_begin, border_size;  // vertices [border_begin, +border_size) of the concatenated border array
  int pad0, pad1;
};
// edges: 3 float4 per border vertex; rays: 10 float4 per plane; ray_edges: 10 drawn edge indices per plane.
void launch_reabsorb_prepare(const float4* border, const ReabsorbPlane* planes, int n_planes, const int32_t* ray_edges,
                             float4* edges, float4* rays, cudaStream_t s);
// R1: appends (point, plane) for every pair with dist <= t; *counter ends at the number of pairs found even when
// it exceeds cap (then nothing beyond cap was written and the caller retries with a larger buffer).
void launch_reabsorb_filter(CloudView cloud, size_t n, const ReabsorbPlane* planes, int n_planes, float t,
                            unsigned long long* counter, uint2* cand, unsigned long long cap, int num_sms, cudaStream_t s);
// R2: claimed[i] = 1, key (plane << 32 | i) appended, plane_counts[plane]++ for every pair inside its polygon.
void launch_reabsorb_poly(CloudView cloud, const uint2* cand, unsigned long long n_cand, const ReabsorbPlane* planes,
                          const float4* edges, const float4* rays, uint32_t* claimed, unsigned long long* keys,
                          unsigned long long* n_absorbed, int32_t* plane_counts, int num_sms, cudaStream_t s);
// R3: sorts the keys (per plane ascending point index) and splits them into current / original index lists.
size_t reabsorb_sort_temp_bytes(size_t n);
void launch_reabsorb_lists(const unsigned long long* keys_in, unsigned long long* keys_sorted, size_t n, int key_bits, void* temp,
                           size_t temp_bytes, const int32_t* orig, int32_t* out_cur, int32_t* out_orig, cudaStream_t s);

// Batch: per-cloud inlier count of per-cloud planes is K2 with K = 1; nothing else needed.

// ---- point normals by radius PCA (pr_normals.cu): pcl::NormalEstimationOMP with a radius search -------------------
struct NormalsGrid {
  double lo[3];        // grid origin (lower corner of the cloud's bounding box)
  double inv_h;        // 1 / cell size, cell size = radius * (1 + 1e-6)
  long long dim[3];    // cells per axis
  unsigned long long no_cell;  // dim[0] * dim[1] * dim[2]: the key of non-finite points
};
size_t normals_sort_temp_bytes(size_t n);
// cell keys -> sorted (key, index) pairs in keys[n..2n), idx[n..2n) -> coordinates in that order (x | y | z, n each)
void launch_normals_sort(CloudView cloud, size_t n, const NormalsGrid& g, int key_bits, unsigned long long* keys, uint32_t* idx,
                         void* temp, size_t temp_bytes, float* sorted_xyz, cudaStream_t s);
// out[idx] = (normal_x, normal_y, normal_z, curvature), NaN for non-finite points and points with < 3 neighbours
void launch_normals(const float* sorted_xyz, const unsigned long long* sorted_keys, const uint32_t* sorted_idx, size_t n,
                    const NormalsGrid& g, float r2, double scale, const float vp[3], float4* out, int32_t* n_neighbors,
                    cudaStream_t s);

// clusterFilt: connected components of the radius graph over the same cell-sorted points (lock-free union-find);
// flags[original index] = 1 for the points of components with at most max_small points.  parent, sizes: n uint32 each.
void launch_cluster_flags(const float* sorted_xyz, const unsigned long long* sorted_keys, const uint32_t* sorted_idx, size_t n,
                          const NormalsGrid& g, float r2, uint32_t max_small, uint32_t* parent, uint32_t* sizes, uint32_t* flags,
                          cudaStream_t s);

// ---- per-round exchanges of the point-sharded path through peer memory (pr_p2p.cu) ------------------------------
// Every rank owns a mailbox in its HBM that all peers map with CUDA IPC; peers[r] is rank r's mailbox as seen from
// this process (peers[rank] is the local allocation).
constexpr int kP2PMaxRanks = 8;
struct P2PView {
  unsigned char* peers[kP2PMaxRanks];
  int n_ranks;
  int rank;
};
// The host-free peel loop lets an exchange kernel also run the small step that consumes its result (same CTA, no extra
// launch): computeModel's decision after the counts, the closed-form plane after the moments, the stop rule after the
// totals, the models after the sample points.
struct P2PTail {
  enum { kNone = 0, kReplay, kFinish, kAdvance, kModels };
  int kind = kNone;
  RoundRecord* rec = nullptr;
  const int32_t* good = nullptr;      // kReplay
  const float4* hyps = nullptr;       // kFinish
  const int32_t* triples = nullptr;   // kFinish
  int optimize = 0, scale_exp = 0, n_draws = 0;  // kFinish
  int min_plane = 0;                  // kAdvance
  float4* hyps_out = nullptr;         // kModels
  int32_t* good_out = nullptr;        // kModels
};
// One single-CTA kernel per exchange: store this rank's contribution into every peer's mailbox slot
// (slot_off + parity * buffer_bytes + rank * slot_stride), raise flag[rank] = epoch in every peer's flag array (flag_off,
// kP2PMaxRanks x uint64), wait (bounded; *err = 1 on timeout) for every rank's flag locally, then sum / concatenate in
// rank order.  The epoch is *epoch_ctr + 1 (device counter, advanced by the kernel); parity = epoch & 1 selects one of
// the channel's two buffers.  st: the exchange is skipped (no epoch consumed) once st->stop is set.  wait_ns (optional,
// 2 x uint64): time spent spinning on the peers' flags and the number of exchanges, accumulated.
void launch_p2p_allreduce_i32(const P2PView& v, const int32_t* src, size_t n, size_t slot_off, size_t buffer_bytes, size_t slot_stride,
                              size_t flag_off, unsigned long long* epoch_ctr, int32_t* dst, unsigned* err, cudaStream_t s,
                              RoundState* st = nullptr, unsigned long long* wait_ns = nullptr, const P2PTail* tail = nullptr,
                              unsigned* tickets = nullptr);
void launch_p2p_allreduce_i64(const P2PView& v, const long long* src, size_t n, size_t slot_off, size_t buffer_bytes, size_t slot_stride,
                              size_t flag_off, unsigned long long* epoch_ctr, long long* dst, unsigned* err, cudaStream_t s,
                              RoundState* st = nullptr, unsigned long long* wait_ns = nullptr, const P2PTail* tail = nullptr);
// dst must directly follow the n source values in memory (src = dst - n) when the kAdvance tail is used.
void launch_p2p_allgather_i64(const P2PView& v, const long long* src, size_t n, size_t slot_off, size_t buffer_bytes, size_t slot_stride,
                              size_t flag_off, unsigned long long* epoch_ctr, long long* dst, unsigned* err, cudaStream_t s,
                              RoundState* st = nullptr, unsigned long long* wait_ns = nullptr, const P2PTail* tail = nullptr);
// K1a fused with its exchange: the owner of sample s writes the point's bits into every rank's sample buffer (sp_off);
// dst receives all n_samples entries.  st: shard extent (first, n) from the device state.
void launch_p2p_samples(const P2PView& v, CloudView cloud, long long first, size_t n, const int32_t* triples, int n_samples,
                        size_t sp_off, size_t buffer_bytes, size_t flag_off, unsigned long long* epoch_ctr, int4* dst, unsigned* err,
                        cudaStream_t s, const RoundState* st = nullptr, unsigned long long* wait_ns = nullptr, const P2PTail* tail = nullptr,
                        unsigned* tickets = nullptr);
// tickets (sample points and counts only): two zero-initialised counters per channel; given them, the exchange runs
// over several blocks (pr_p2p.cu), which leave them zero again.

// ---- the peel loop without the host (pr_chain.cu): per-round kernels driven by a RoundState in HBM ------------------
// PCL's index triples for a cloud of st->n_global points: rnd = the first 3 * n_draws values of mt19937(seed) >> 1,
// table = draw_table_slots(n_draws) uint64 slots tagged with a 16-bit epoch (never cleared between rounds: pass a new
// epoch in 1 .. 65535 per call and zero the table once / when the epoch wraps), coll = kDrawCollCap uint32 + coll_count
// (zeroed by the caller before the first round, by the resolve kernel afterwards).  The scatter launch also clears what
// the round accumulates into: the K counts, the refit moments, the compaction descriptors (scratch) and the
// blocks-done ticket.  Sets st->stop = 2 when the round has to go back to the sequential host sampler, 1 when the cloud
// has fewer than 3 points.
size_t draw_table_slots(int n_draws);
void launch_draw(const uint32_t* rnd, int n_draws, RoundState* st, int32_t* triples, unsigned long long* table, size_t table_slots,
                 uint32_t epoch, uint32_t* coll, uint32_t* coll_count, RoundRecord* rec, int32_t* counts, RefitOut* refit, void* scratch,
                 size_t scratch_bytes, unsigned* tickets, cudaStream_t s);
// computeModel's decision over K counts (score-all mode): st->best / best_count, or st->stop = 2 when a bad sample means
// PCL would draw beyond the K scored hypotheses.
void launch_replay(const int32_t* counts, const int32_t* good, int K, RoundState* st, RoundRecord* rec, cudaStream_t s);
// st->plane = refined coefficients (closed form from the summed moments, pr_math.h) or the raw model; fills the record.
void launch_finish(RoundState* st, const float4* hyps, const int32_t* triples, const RefitOut* refit, int optimize, int scale_exp,
                   int n_draws, RoundRecord* rec, cudaStream_t s);

// ---- batch of small clouds without the host in the loop (score-all mode; pr_chain.cu) ----------------------------------
// K1 for a batch whose clouds all draw the same K triples: sample points + models per (cloud, draw); clears counts and flag.
void launch_batch_gather_models(CloudView clouds, size_t n_per, size_t stride, int n_clouds, const int32_t* triples, int K, int4* sample_pts,
                                float4* hyps, int32_t* good, int32_t* counts, int* flag, cudaStream_t s);
// best[c] = computeModel's winner among cloud c's K draws (-1 + *any_bad when a degenerate sample needs PCL's redraw).
// refit_to_clear (optional): cloud c's moments are zeroed on the way (the refit pass accumulates into them).
void launch_batch_replay(const int32_t* counts, const int32_t* good, int K, int n_clouds, int32_t* best, int32_t* best_count, int* any_bad,
                         RefitOut* refit_to_clear, cudaStream_t s);
// raw[c] = hyps[c * K + best[c]]; refined[c] = closed-form plane from refit[c] on the 2^-scale_exp[c] grid (or raw[c]).
void launch_batch_finish(const float4* hyps, int K, const int32_t* best, const RefitOut* refit, const int32_t* scale_exp, int optimize,
                         int n_clouds, float4* raw, float4* refined, cudaStream_t s);
// cnt[c] = |{ i in cloud c : |planes[c] . (p_i, 1)| < t }| (0 where best[c] < 0)
void launch_batch_count(CloudView clouds, size_t n_per, size_t stride, int n_clouds, const float4* planes, const int32_t* best, float t,
                        int dot_order, int32_t* cnt, cudaStream_t s);
// offs = exclusive scan of the per-cloud final counts; out (optional, cap entries) receives every cloud's ascending
// inlier indices at offs[c] (clouds whose list would not fit are skipped).
void launch_batch_lists(CloudView clouds, size_t n_per, size_t stride, int n_clouds, const float4* planes, const int32_t* best, float t,
                        int dot_order, const int32_t* cnt, unsigned long long* offs, size_t cap, int32_t* out, cudaStream_t s);

// Measurement helpers.
void launch_ffma_peak(float* out, int iters, int grid, cudaStream_t s);
void launch_copy(const float4* src, float4* dst, size_t n_vec, int num_sms, cudaStream_t s);
void launch_fill(float4* dst, size_t n_vec, float v, int num_sms, cudaStream_t s);

}  // namespace pr
