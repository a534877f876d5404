/*
 * plane_ransac.h — C ABI of the B200 (sm_100a) plane-detection backend.
 *
 * Drop-in boundary for ONE path of czh55/Dialog: multi-plane extraction from a
 * pcl::PointCloud<pcl::PointXYZ> by RANSAC plane segmentation with iterative inlier peeling
 * (PCL 1.8 SACSegmentation<PointXYZ>(SACMODEL_PLANE, SAC_RANSAC) + ExtractIndices semantics).
 * The reference has no FFI layer of its own; every entry point below names the
 * reference (or PCL 1.8) interface it stands in for.
 *
 *   pr_point                       pcl::PointXYZ / PointT            Dialog/HeaderFile.h:53-54
 *   pr_params.distance_threshold   T_dist_point_plane                Dialog/PlaneDetect.h:88, config.txt:29
 *   pr_params.min_plane_size       T_num_of_single_plane             Dialog/PlaneDetect.h:70, config.txt:20
 *   pr_params.max_iterations       SACSegmentation::setMaxIterations (no counterpart in config.txt)
 *   plane coefficients (a,b,c,d)   Plane::coeff.values               Dialog/HeaderFile.h:81-88, PlaneDetect.h:1493-1497
 *   inlier indices                 Plane::points_set                 Dialog/HeaderFile.h:85
 *   remaining cloud                rebuilt source_cloud              Dialog/PlaneDetect.h:1560-1566
 *
 * All types are POD; no C++ types or exceptions cross the boundary.  Every call returns 0 on
 * success or a negative pr_status; plane_ransac_last_error() holds the message of the calling
 * thread's last failure.  One context owns one CUDA device + stream and is not thread-safe (the
 * reference drives this path from the GUI thread only, Dialog/PCLViewer.cpp:1180-1235).
 * There is no CPU fallback: without a CUDA device plane_ransac_create fails.
 */
#ifndef PLANE_RANSAC_H
#define PLANE_RANSAC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLANE_RANSAC_ABI_VERSION 3

typedef struct plane_ransac_ctx plane_ransac_ctx;

/* == pcl::PointXYZ: 16 bytes, w is padding (treated as 1). */
typedef struct { float x, y, z, w; } pr_point;

typedef enum {
  PR_OK = 0,
  PR_ERR_INVALID = -1,   /* bad argument                                              */
  PR_ERR_CUDA = -2,      /* CUDA runtime / driver error                               */
  PR_ERR_NO_CLOUD = -3,  /* no cloud staged                                           */
  PR_ERR_CAPACITY = -4,  /* caller buffer too small                                   */
  PR_ERR_COMM = -5,      /* NCCL error or NCCL not loadable                           */
  PR_ERR_OOM = -6        /* host or device allocation failed                          */
} pr_status;

/* Order of the FP32 dot product coeff·(x,y,z,1) in the distance test.
 *   PR_DOT_PCL_SSE2  (a*x + c*z) + (b*y + d), products and sums rounded separately: Eigen's SSE2
 *                    packet reduction as PCL 1.8 / MSVC v140 runs it.  6 FP32 ops per point-hypothesis.
 *   PR_DOT_FMA       fma(a,x, fma(b,y, fma(c,z,d))): 3 FFMA per point-hypothesis (default).
 * The two differ only for points whose residual is within rounding of the threshold. */
enum { PR_DOT_PCL_SSE2 = 0, PR_DOT_FMA = 1 };

/* How countWithinDistance is evaluated for the batch of hypotheses.  Both give identical counts.
 *   PR_SCORER_BRUTE  every hypothesis against every point (the FP32-FMA-bound kernel; default).
 *   PR_SCORER_HIER   a Morton-sorted copy of the cloud with a bounding box per 32 points: blocks whose box lies
 *                    outside (or entirely inside) the threshold slab of a hypothesis are skipped (or counted
 *                    whole); the rest are evaluated point by point with the same arithmetic.  Costs one sort per
 *                    staged cloud and a second compaction per peel round; not used by the batch entry points. */
enum { PR_SCORER_BRUTE = 0, PR_SCORER_HIER = 1 };

/* How optimizeModelCoefficients (the least-squares refit behind pcl::computePointNormal, Dialog/PlaneDetect.h:1084,
 * 1130,1386,1485) accumulates the covariance.
 *   PR_REFIT_FIXED      (default) inlier coordinates quantised to a 2^-30-of-the-extent grid, moments summed as exact
 *                       integers, pcl::eigen33 in double: independent of any summation order, so 1 thread, 148 SMs and
 *                       8 GPUs give the same bits.
 *   PR_REFIT_PCL_FLOAT  PCL 1.8's own arithmetic: nine FP32 accumulators summed sequentially over the inliers in index
 *                       order (computeMeanAndCovarianceMatrix), pcl::eigen33 in FP32.  One device thread does the sums
 *                       (about 20 ms per million inliers), the host the 3x3 solve; a parity mode, host-driven rounds, one
 *                       GPU only (a sequential sum over a sharded cloud would need the points in one place).
 * The two agree to about 1e-5 relative in the coefficients (tests/test_gpu_parity.py enumerates the inliers that differ). */
enum { PR_REFIT_FIXED = 0, PR_REFIT_PCL_FLOAT = 1 };

/* pcl::SACSegmentation knobs + the peel stop rule. */
typedef struct {
  double distance_threshold;  /* setDistanceThreshold(double); inlier iff |n·p + d| <  t (strict) */
  int max_iterations;         /* setMaxIterations; up to max_iterations + 1 trials are scored    */
  int min_plane_size;         /* stop peeling when a plane has fewer inliers                     */
  double probability;         /* setProbability; 0.99 = PCL default; 1.0 = score every trial     */
  int optimize_coefficients;  /* setOptimizeCoefficients: least-squares refit + re-selection     */
  unsigned seed;              /* 12345u = PCL's non-random SampleConsensusModel seed             */
  int max_planes;             /* bound on planes returned by plane_ransac_extract_planes         */
  int dot_order;              /* PR_DOT_*                                                         */
  int scorer;                 /* PR_SCORER_*: how the inlier counts are computed (same counts either way) */
  int refit_mode;             /* PR_REFIT_* (ABI 3)                                                */
} pr_params;

/* What one segment() call decided (mirrors RandomSampleConsensus state; used by parity tests). */
typedef struct {
  int ok;               /* a model was found                                                    */
  int iterations;       /* RandomSampleConsensus::iterations_                                   */
  int draws;            /* drawIndexSample calls consumed                                       */
  int skipped;          /* skipped_count                                                        */
  int best_sample[3];   /* model_ (indices into the cloud of this round, global when sharded)   */
  int best_count;       /* inliers of the raw model                                             */
  float raw_coeff[4];   /* model_coefficients_ before the refit                                 */
  int n_inliers_raw;    /* == best_count                                                        */
  int n_inliers;        /* final inlier count (after refit + re-selection), global when sharded */
  int scale_exp;        /* refit grid exponent s (grid = 2^-s)                                  */
  int n_scored;         /* hypotheses scored on the device for this call                        */
  long long n_cloud;    /* points in the cloud of this round (global when sharded)              */
} pr_segment_info;

#define PR_LOOP_STAGES 9
/* Device time per kernel class, accumulated while profiling is enabled (CUDA events on the
 * context's stream), and launch counts since the last plane_ransac_profile_reset. */
typedef struct {
  double ms_stage, ms_models, ms_score, ms_refit, ms_compact, ms_other;
  long long launches_stage, launches_models, launches_score, launches_refit, launches_compact,
      launches_other;
  long long pairs_scored;    /* sum of points x hypotheses over score launches (this rank)      */
  long long points_refit;    /* points streamed by refit launches                               */
  long long points_compact;  /* points streamed by compaction launches                          */
  long long bytes_compact;   /* bytes compaction launches move: 12 B per point read (16 B when the cloud carries an
                                original-index plane, i.e. after the first peel), 16 B per kept point, 4 B per
                                inlier and list written                                           */
  long long bytes_refit;     /* algorithmic bytes of refit launches                             */
  /* host wall time inside segment/extract calls, by phase (always accumulated) */
  double host_ms_sampling;   /* drawing PCL's index triples                                     */
  double host_ms_replay;     /* replaying RANSAC's sequential decisions + plane from moments    */
  double host_ms_wait;       /* blocked in stream synchronisation (device work + copies)        */
  double host_ms_total;      /* whole calls                                                     */
  /* peel launches of segment / extract calls (ABI 3): what the compaction kept and what it peeled off */
  long long points_kept;     /* points written to the remaining cloud                           */
  long long points_peeled;   /* inliers written to the index lists                              */
  /* sharded runs with peer-memory exchanges: per channel (0 sample points, 1 counts, 2 refit moments, 3 remaining
   * counts) the time this rank's exchange kernels spent spinning on the peers' flags — the slowest rank's lag plus the
   * NVLink round trip — and the number of exchanges, since the last profile_reset */
  double p2p_wait_ms[4];
  long long p2p_exchanges[4];
  /* device-resident round loop: time from the start of one kernel of a round to the start of the next (the kernel plus
   * the hand-over), from %globaltimer stamps the kernels take themselves — no events, always accumulated, also while
   * profiling is off.  [0] sampler, parallel phase  [1] sampler, collision replay  [2] sample points (+ exchange) and
   * models  [3] scoring  [4] (count exchange +) computeModel's decision  [5] refit moments  [6] closed-form plane as its
   * own step (moment exchange / no refit pass)  [7] peel  [8] remaining-count exchange + stop rule (sharded).  Their
   * sum is the device time of loop_rounds rounds. */
  double loop_ms[PR_LOOP_STAGES];
  long long loop_rounds;
} pr_profile;

/* ---- lifetime ---------------------------------------------------------------------------- */
int plane_ransac_abi_version(void);
const char* plane_ransac_last_error(void);
void plane_ransac_default_params(pr_params* p); /* t=0.1, it=50, min=500, p=0.99, refit, 12345, 64, FMA, brute */
int plane_ransac_create(plane_ransac_ctx** ctx, int device_id);
void plane_ransac_destroy(plane_ransac_ctx* ctx);

/* ---- staging: replaces preProcess's copy of source_cloud (Dialog/PlaneDetect.h:449-455) ----
 * Uploads the AoS cloud and transposes it to x[] y[] z[] planes in HBM (128-bit aligned, padded
 * with NaN to a tile multiple).  The staged cloud is immutable; extract/segment calls start from it
 * (plane_ransac_extract_planes may be called repeatedly).  pts may be pageable or pinned host memory. */
int plane_ransac_set_cloud(plane_ransac_ctx* ctx, const pr_point* pts, size_t n);
/* Same, from a device pointer (AoS pr_point[] already in HBM on the context's device). */
int plane_ransac_set_cloud_device(plane_ransac_ctx* ctx, const pr_point* dev_pts, size_t n);
/* Like plane_ransac_set_cloud, but returns as soon as the upload is queued: the cloud travels in chunks on a copy
 * stream and the next plane_ransac_extract_planes / plane_ransac_segment_one call stages and scores each chunk as it
 * lands (the sample points of its first batch are read from pts directly), so the host-to-device copy hides under
 * the first scoring pass.  pts must be page-locked (plane_ransac_host_alloc, cudaHostAlloc, a pinned torch tensor) and
 * must stay valid and unchanged until the next call on ctx returns; any other call first waits for the upload.
 * Pageable memory and clouds under 2M points take the synchronous path of plane_ransac_set_cloud (in a sharded context
 * every rank must pass the same kind of buffer and a shard of the same size class, so that all make the same choice). */
int plane_ransac_set_cloud_async(plane_ransac_ctx* ctx, const pr_point* pts, size_t n);
/* Staging with the reference's preProcess steps fused in (Dialog/PlaneDetect.h:449-481):
 *   PR_STAGE_REMOVE_NONFINITE     pcl::removeNaNFromPointCloud: points with a NaN/Inf coordinate are dropped,
 *                                 order preserved; later indices refer to the filtered cloud.
 *   PR_STAGE_TRANSLATE_CENTROID   isConductTranslate: subtract the centroid of the (filtered) cloud, one float
 *                                 subtraction per coordinate.  The centroid is the exactly rounded mean (integer
 *                                 sums on the refit grid), not the reference's sequential float sum, so that
 *                                 any thread/GPU count gives the same bits.
 * n_kept / centroid are optional outputs.  plane_ransac_staged_source_indices returns, for each staged point,
 * its index in the array given to set_cloud_ex (identity without the filter). */
enum { PR_STAGE_REMOVE_NONFINITE = 1, PR_STAGE_TRANSLATE_CENTROID = 2 };
int plane_ransac_set_cloud_ex(plane_ransac_ctx* ctx, const pr_point* pts, size_t n, unsigned flags, size_t* n_kept,
                              float centroid[3]);
int plane_ransac_staged_source_indices(plane_ransac_ctx* ctx, int32_t* out, size_t cap);
/* n_staged: points given to set_cloud; n_current: points left after the last extract call. */
int plane_ransac_cloud_size(plane_ransac_ctx* ctx, size_t* n_staged, size_t* n_current);

/* ---- countWithinDistance for K models (bench + parity hook) --------------------------------
 * triples: 3*K indices into the staged cloud (as drawn by SampleConsensusModel::drawIndexSample).
 * counts[k] = |{ i : |coeff_k · (p_i,1)| < t }|; coeffs (optional, 4*K) and good (optional, K)
 * receive computeModelCoefficients' output and isSampleGood's verdict; bad samples count 0. */
int plane_ransac_score(plane_ransac_ctx* ctx, const int32_t* triples, int K, double t, int dot_order,
                       int32_t* counts, float* coeffs, uint8_t* good);

/* ---- pcl::SACSegmentation<PointXYZ>::segment(inliers, coefficients) on the staged cloud -----
 * inliers: ascending indices, capacity cap.  On "no model" n_inliers = 0 and coeff is zeroed. */
int plane_ransac_segment_one(plane_ransac_ctx* ctx, const pr_params* prm, float coeff[4], int32_t* inliers,
                             size_t cap, size_t* n_inliers, pr_segment_info* info);

/* ---- segment + ExtractIndices(negative) peel loop -------------------------------------------
 * coeffs: 4*max_planes floats.  inlier_cur (optional): per plane, indices into the cloud of that
 * round (what PCL's loop yields); inlier_orig (optional): the same points as indices into the staged
 * cloud; both have capacity idx_cap entries in total and are delimited by plane_offsets
 * (max_planes + 1 entries).  infos (optional): max_planes + 1 entries, the last rejected segment
 * call included.  When sharded, indices are local to this rank's shard. */
int plane_ransac_extract_planes(plane_ransac_ctx* ctx, const pr_params* prm, float* coeffs,
                                int32_t* inlier_cur, int32_t* inlier_orig, size_t idx_cap,
                                size_t* plane_offsets, int* n_planes, pr_segment_info* infos);
/* Who drives the peel loop of plane_ransac_extract_planes.  Both give identical results.
 *   PR_LOOP_AUTO  (default) in score-all mode (probability = 1, brute scorer, at most 16384 hypotheses per round) whole
 *                 rounds are queued on the device — PCL's index triples, computeModel's decision, the closed-form
 *                 refit and the minimum-size rule run as kernels on a device-resident round state — and the host only
 *                 reads each round's record as it completes; rounds that need PCL's redraw rule fall back to the host.
 *   PR_LOOP_HOST  every round is driven by the host: it draws the triples, replays computeModel over the counts and
 *                 solves the refit, with three stream synchronisations per round (the adaptive mode always runs so). */
enum { PR_LOOP_AUTO = 0, PR_LOOP_HOST = 1 };
int plane_ransac_set_round_loop(plane_ransac_ctx* ctx, int mode);
/* Plane::points_set of plane k of the last extract call (Dialog/HeaderFile.h:85): its inlier points in index
 * order, read from the staged cloud.  With project != 0 every point is projected onto the plane with the
 * reference's projPoint2Plane arithmetic (Dialog/PlaneDetect.h:1442-1448) — the cloud polyPointCloud hands to
 * pcl::ConcaveHull (Dialog/PlaneDetect.h:1391-1401).  out may be NULL to query the count. */
int plane_ransac_plane_points(plane_ransac_ctx* ctx, int plane_index, int project, pr_point* out, size_t cap, size_t* n);
/* Points left after the last extract call, original order (== the rebuilt source_cloud). */
int plane_ransac_remaining(plane_ransac_ctx* ctx, pr_point* out, size_t cap, size_t* n);

/* ---- postProcessPlanes re-absorption (Dialog/PlaneDetect.h:1454-1580) ------------------------------------------
 * The reference's own use of the threshold test + peel: every point of the current cloud (what is left after the last
 * extract call, or the staged cloud) is tested against every plane polygon with isPointInPoly (:1891-1955): distance to
 * the plane <= dist_threshold (T_dist_point_plane, non-strict) by projPoint2Plane / distP2P, then ten in-plane rays,
 * each perpendicular to a border edge drawn with rand() % border.size() after srand(rand_seed) — the reference seeds
 * with time(0) at every call — intersected with every border edge by isBothLineSegsIntersect (:1957-2016); inside when
 * at least five rays cross an odd number of edges.  A point joins every plane that contains it (:1549-1555 has no
 * break) and the unclaimed points become the current cloud in their original order (:1560-1566), which
 * plane_ransac_remaining returns.
 *   coeffs: 4 floats per plane (a, b, c, d).  border: the polygons' vertices (Plane::border, e.g. from
 *   pcl::ConcaveHull) concatenated; plane j owns border[border_offsets[j] .. border_offsets[j + 1]) (>= 1 vertex).
 *   absorbed_cur / absorbed_orig (optional, idx_cap entries each): per plane the ascending indices of the points it
 *   claimed, as positions in the cloud before this call / as indices into the staged cloud, delimited by
 *   plane_offsets (n_planes + 1 entries).  When sharded every rank processes its shard; indices are rank-local. */
int plane_ransac_reabsorb(plane_ransac_ctx* ctx, const float* coeffs, const pr_point* border, const size_t* border_offsets,
                          int n_planes, float dist_threshold, unsigned rand_seed, int32_t* absorbed_cur,
                          int32_t* absorbed_orig, size_t idx_cap, size_t* plane_offsets, size_t* n_remaining);

/* ---- estimateNormal() (Dialog/PlaneDetect.h:515-545): pcl::NormalEstimationOMP<PointXYZ, Normal> with
 * setRadiusSearch(radius) (r_for_estimate_normal, Dialog/config.txt:4) on the current cloud.  Neighbours of a point: the
 * finite points whose FP32 squared distance (FLANN's L2_Simple order) is strictly below (float)(radius * radius), the
 * point itself included; fewer than three, or a non-finite point, give NaN.  normal = eigenvector of the smallest
 * eigenvalue of the neighbourhood covariance (PCL's eigen33 closed form), curvature = |lambda_0 / trace|, flipped
 * towards viewpoint (NULL = the origin, PCL's default).  The covariance is accumulated as exact integers on a grid of
 * 2^-18 of the radius about each point, so the result does not depend on any traversal order (PCL's own float sums
 * depend on FLANN's neighbour order; they agree to ~1e-3 in the normal on noisy data).  n_neighbors is optional.
 * Single GPU only. */
typedef struct { float normal_x, normal_y, normal_z, curvature; } pr_normal;   /* == pcl::Normal's first 3 + curvature */
int plane_ransac_estimate_normals(plane_ransac_ctx* ctx, double radius, const float viewpoint[3], pr_normal* out,
                                  size_t cap, int32_t* n_neighbors);

/* ---- clusterFilt() (Dialog/PlaneDetect.h:1582-1656), the last step of postProcessPlanes: the current cloud is split into
 * the connected components of its radius graph (an edge where FLANN's FP32 squared distance is < (float)(radius * radius);
 * the reference uses radius_local, Dialog/config.txt:22) and every component with at most max_small_cluster points
 * (T_cluster_num, config.txt:30; the reference tests "size <= T") is dropped; the rest stays in its original order.
 * The reference grows clusters by BFS; components do not depend on the order, so a parallel union-find gives the same
 * result.  (One deviation: the reference skips the first radius-search hit as "the point itself", which splits exact
 * duplicates off when FLANN lists the duplicate first; here duplicates are connected.)  Single GPU only. */
int plane_ransac_cluster_filter(plane_ransac_ctx* ctx, double radius, int max_small_cluster, size_t* n_removed,
                                size_t* n_remaining);

/* ---- "run again" (PCLViewer::on_runAgainAction_triggered, Dialog/PCLViewer.cpp:1120-1178) ------------------------
 * The reference reruns its pipeline on the shrunken source_cloud that postProcessPlanes left (Dialog/PlaneDetect.h:
 * 1560-1572).  This makes the current cloud (what the last extract / reabsorb call left) the staged cloud, on the
 * device, so the next segment / extract / score call works on it; plane_ransac_staged_source_indices then maps its
 * points to the caller's original array.  plane_ransac_plane_points of earlier planes is no longer available. */
int plane_ransac_restage_remaining(plane_ransac_ctx* ctx);

/* ---- batch of equal-sized small clouds (per-scan tiles), one best plane each, no peel --------
 * pts: n_clouds * n_per_cloud points.  Every cloud runs segment() with the same parameters (and,
 * having the same size and seed, the same index triples).  coeffs: 4*n_clouds; n_inliers: n_clouds.
 * Single best plane per cloud, no peel (a tile that needs its planes peeled is a cloud for plane_ransac_extract_planes). */
int plane_ransac_set_cloud_batch(plane_ransac_ctx* ctx, const pr_point* pts, size_t n_clouds,
                                 size_t n_per_cloud);
int plane_ransac_segment_batch(plane_ransac_ctx* ctx, const pr_params* prm, float* coeffs,
                               int32_t* n_inliers, pr_segment_info* infos /* optional, n_clouds */);
/* The same with segment()'s `inliers` output for every cloud: offsets (n_clouds + 1 entries) delimits the clouds' lists
 * inside `inliers` (capacity cap entries in total; NULL to get the offsets alone); each list holds ascending indices into
 * its own cloud.  When the lists do not fit, coefficients, counts and offsets are still filled and PR_ERR_CAPACITY is
 * returned (offsets[n_clouds] is the capacity needed).  In score-all mode the whole batch runs without the host in the
 * loop (see PR_LOOP_AUTO); a shard of a larger batch (cloud_id % n_gpu == rank) is just a smaller batch: no collective. */
int plane_ransac_segment_batch_lists(plane_ransac_ctx* ctx, const pr_params* prm, float* coeffs, int32_t* n_inliers,
                                     int32_t* inliers, size_t cap, size_t* offsets, pr_segment_info* infos);

/* ---- point-sharded multi-GPU (one process per GPU; NCCL over NVLink) -------------------------
 * Rank r stages its contiguous index range of the global cloud with plane_ransac_set_cloud; after
 * comm_init every segment/extract/score call is collective: per-hypothesis counts and refit moments
 * are summed with ncclAllReduce (integers, so the result is bit-identical to one GPU). */
#define PLANE_RANSAC_UNIQUE_ID_BYTES 128
int plane_ransac_comm_unique_id(void* out128);
int plane_ransac_comm_init(plane_ransac_ctx* ctx, int n_ranks, int rank, const void* unique_id128);
/* 1 when the per-round exchanges (sample points, per-hypothesis counts, refit moments, remaining counts) run as
 * peer-memory kernels: every rank owns a mailbox in its HBM that the peers map with CUDA IPC and write over NVLink,
 * flags with system-scope release/acquire, sums in rank order (bit-identical to one GPU).  0: NCCL collectives (set
 * PR_P2P=0, more than 8 ranks, or a peer that could not be mapped — decided collectively at comm_init). */
int plane_ransac_comm_p2p_enabled(plane_ransac_ctx* ctx);
/* Global size and this rank's first global index for the staged / current cloud. */
int plane_ransac_shard_info(plane_ransac_ctx* ctx, long long* n_global_staged, long long* first_staged,
                            long long* n_global_current, long long* first_current);

/* ---- pinned host buffers ---------------------------------------------------------------------
 * Page-locked host memory for clouds and index lists (cudaMallocHost / cudaFreeHost): copies from and to
 * such buffers run at full PCIe rate.  Any host pointer is accepted by the calls above; these are a
 * convenience for callers without a CUDA runtime of their own. */
int plane_ransac_host_alloc(size_t bytes, void** out);
int plane_ransac_host_free(void* p);
/* Page-locks memory the caller already owns (cudaHostRegister / cudaHostUnregister), e.g. the storage of a
 * pcl::PointCloud (the reference's source_cloud, Dialog/PlaneDetect.h:104), so that plane_ransac_set_cloud_async takes
 * the overlapped upload on it.  Unregister before the memory is freed or reallocated. */
int plane_ransac_host_register(void* p, size_t bytes);
int plane_ransac_host_unregister(void* p);

/* ---- cloud reader: pcl::io::loadPCDFile for an XYZ cloud (Dialog/PCLViewer.cpp:80-89) ---------------------------
 * PCD v0.7, DATA ascii / binary / binary_compressed; x, y, z may sit anywhere in the record and have any numeric
 * type, other fields are skipped (Dialog/double_shadow.pcd is "x y z rgb", ascii).  *points is a page-locked buffer of
 * *n_points pcl::PointXYZ-compatible records (w = 1), ready for plane_ransac_set_cloud_async; release it with
 * plane_ransac_host_free.  Non-finite points are kept (plane_ransac_set_cloud_ex removes them, as the reference's
 * preProcess does). */
int plane_ransac_load_pcd(const char* path, pr_point** points, size_t* n_points);

/* ---- measurement ------------------------------------------------------------------------- */
int plane_ransac_profile_enable(plane_ransac_ctx* ctx, int on);
int plane_ransac_profile_reset(plane_ransac_ctx* ctx);
int plane_ransac_profile_get(plane_ransac_ctx* ctx, pr_profile* out);
/* Rounds the last plane_ransac_extract_planes call ran in the device-resident loop, in order: PR_LOOP_STAGES + 1
 * %globaltimer values (ns) per round — when each stage of pr_profile.loop_ms started (0: the round has no such stage) and
 * when the round's record was written.  stamps may be NULL (count only); at most cap_rounds rounds are copied. */
int plane_ransac_round_timeline(plane_ransac_ctx* ctx, unsigned long long* stamps, size_t cap_rounds, size_t* n_rounds);
/* Device-side stopwatch: CUDA events recorded on the context's stream (the stream every kernel and
 * copy of this library is issued on).  stop synchronises and returns the elapsed milliseconds. */
int plane_ransac_timer_start(plane_ransac_ctx* ctx);
int plane_ransac_timer_stop(plane_ransac_ctx* ctx, double* ms);
/* FP32 FMA peak of this device measured with an FFMA-only kernel (TFLOP/s), for the scoring roofline. */
int plane_ransac_measure_ffma_peak(plane_ransac_ctx* ctx, double* tflops);
/* Streaming copy bandwidth of this device (GB/s, read + write bytes), for the HBM rooflines. */
int plane_ransac_measure_copy_bw(plane_ransac_ctx* ctx, size_t bytes, double* gbs);
/* Writes a buffer larger than L2 (between timed iterations). */
int plane_ransac_flush_l2(plane_ransac_ctx* ctx);

/* ---- host-side logic, exported for tests and multi-process drivers (no device work) -----------
 * The first n_draws triples SampleConsensusModel::drawIndexSample yields for n points. */
int plane_ransac_host_draw_triples(size_t n_points, unsigned seed, int n_draws, int32_t* triples);
/* The same triples through the parallel formulation the device-side round loop uses (csrc/pr_draw.h): every op of the
 * partial Fisher-Yates walk is evaluated independently and only the ops whose position was also picked by another op
 * are replayed sequentially.  *fell_back = 1 (triples untouched beyond scratch use) when more ops collide than the
 * device replays; the device then leaves that round to the sequential host sampler. */
int plane_ransac_host_draw_triples_parallel(size_t n_points, unsigned seed, int n_draws, int32_t* triples, int* fell_back);
/* RandomSampleConsensus::computeModel's loop replayed over per-draw results: counts[j] is the
 * inlier count of draw j, good[j] isSampleGood's verdict.  Returns the index of the winning draw in
 * *best_draw (-1: none), fills iterations/draws_used/skipped, and sets *exhausted when the loop
 * would have needed more than n_draws draws. */
int plane_ransac_host_replay(const int32_t* counts, const uint8_t* good, int n_draws, long long n_points,
                             int max_iterations, double probability, int* best_draw, int* iterations,
                             int* draws_used, int* skipped, int* exhausted);
/* The ten border-edge indices isPointInPoly draws: srand(seed), then rand() % border_size with the MSVC CRT
 * generator (the reference is built with MSVC v140: Dialog/Dialog.vcxproj:20). */
int plane_ransac_host_rand_edges(unsigned seed, int border_size, int32_t edges[10]);
/* Contiguous shard [first, first + count) of rank r among n_ranks for n points. */
int plane_ransac_host_shard_range(long long n_points, int n_ranks, int rank, long long* first,
                                  long long* count);
/* Least-squares plane from exact integer moments (see DESIGN.md "refit"): m = {n, Sx, Sy, Sz,
 * Sxx_hi, Sxx_lo, ... Szz_hi, Szz_lo}, S_ab = hi * 2^32 + lo. */
int plane_ransac_host_plane_from_moments(const int64_t m[16], const float pivot[3], int scale_exp,
                                         float coeff[4]);

/* PR_REFIT_PCL_FLOAT's host half: PCL 1.8 computeMeanAndCovarianceMatrix's division + covariance, pcl::eigen33 in FP32 and
 * the Hessian d from the nine sequential sums (xx, xy, xz, yy, yz, zz, x, y, z) over n_inliers points. */
int plane_ransac_host_plane_from_pcl_float_sums(const float sums[9], long long n_inliers, float coeff[4]);

#ifdef __cplusplus
}
#endif
#endif /* PLANE_RANSAC_H */
