/*
 * pr_oracle.h — CPU oracle for the plane-RANSAC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (dialog_b200/, include/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the path named by BASELINE.json (pcl::SACSegmentation<PointXYZ> with
 * SACMODEL_PLANE / SAC_RANSAC, iterated with pcl::ExtractIndices) is PCL 1.8 library behaviour.
 * PCL is an un-vendored dependency of the reference (Dialog/PropertySheet-success.props:6,11 pins
 * "pcl-1.8" only), it is absent from /root/reference and from this image, and the reference holds no
 * tests, golden vectors or recorded outputs for it.  This file restates the
 * published PCL 1.8 algorithm; every function names the PCL source it follows.  What IS pinned:
 * the mt19937 stream (C++11 [rand.predef] 10000th-value check and numpy's MT19937), analytic
 * known-answer cases, and the reference's only fixture for this path (Dialog/double_shadow.pcd).
 *
 * Reference-side sites that compute the same quantities (for parity reading):
 *   point-to-plane threshold test  Dialog/PlaneDetect.h:1442-1448, 2019-2023, 1902
 *   peel / order-preserving compaction  Dialog/PlaneDetect.h:1560-1566
 *   least-squares plane refit (pcl::computePointNormal)  Dialog/PlaneDetect.h:1084,1130,1386,1485
 *   parameters T_dist_point_plane / T_num_of_single_plane  Dialog/config.txt:29,20
 */
#ifndef PR_ORACLE_H
#define PR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == pcl::PointXYZ (16 bytes; Dialog/HeaderFile.h:53-54).  w is padding and is treated as 1. */
typedef struct { float x, y, z, w; } orc_point;

/* Order in which the 4-term FP32 dot product coeff·(x,y,z,1) is evaluated.
 *   ORC_DOT_PCL_SSE2: (a*x + c*z) + (b*y + d), every product and sum rounded separately — Eigen's
 *                     4-wide SSE2 packet product + predux (movehl/add, shuffle/add) as PCL 1.8 built
 *                     with MSVC v140 executes it.
 *   ORC_DOT_FMA:      fma(a, x, fma(b, y, fma(c, z, d))) — three fused multiply-adds, the order the
 *                     CUDA scoring kernel uses in its default mode. */
enum { ORC_DOT_PCL_SSE2 = 0, ORC_DOT_FMA = 1 };

/* How optimizeModelCoefficients accumulates the moments.
 *   ORC_REFIT_PCL_FLOAT: PCL 1.8 computeMeanAndCovarianceMatrix — nine float accumulators summed
 *                        sequentially, cov = E[ab] - E[a]E[b], float eigen33.
 *   ORC_REFIT_FIXED:     order-independent exact integer moments of the coordinates quantised to a
 *                        2^-30 grid of the cloud extent about a pivot, eigen33 in double.  This is
 *                        the arithmetic a parallel device can reproduce bit for bit. */
enum { ORC_REFIT_PCL_FLOAT = 0, ORC_REFIT_FIXED = 1 };

typedef struct {
  double distance_threshold;  /* SACSegmentation::setDistanceThreshold(double)              */
  int max_iterations;         /* SACSegmentation::setMaxIterations (PCL default 50)           */
  int min_plane_size;         /* peel stop rule; reference analogue T_num_of_single_plane    */
  double probability;         /* SACSegmentation::setProbability (default 0.99); 1.0 = no early exit */
  int optimize_coefficients;  /* SACSegmentation::setOptimizeCoefficients (default true)      */
  unsigned seed;              /* 12345u == PCL's non-random model seed                        */
  int max_planes;             /* peel loop bound                                              */
  int dot_order;              /* ORC_DOT_*                                                    */
  int refit_mode;             /* ORC_REFIT_*                                                  */
} orc_params;

/* Everything segment() decided, for parity tests. */
typedef struct {
  int ok;                /* computeModel returned true                                   */
  int iterations;        /* RandomSampleConsensus::iterations_ at exit                   */
  int draws;             /* drawIndexSample calls made (good + rejected)                 */
  int skipped;           /* skipped_count                                                */
  int best_sample[3];    /* model_                                                       */
  int best_count;        /* n_best_inliers_count                                         */
  float raw_coeff[4];    /* model_coefficients_ before optimizeModelCoefficients         */
  int n_inliers_raw;     /* |selectWithinDistance(raw_coeff)|                            */
  int n_inliers;         /* final inlier count                                           */
  int scale_exp;         /* ORC_REFIT_FIXED: s (grid = 2^-s)                             */
} orc_trace;

/* ---- mt19937 (boost::mt19937 == std::mt19937) ------------------------------------------- */
typedef struct { uint32_t mt[624]; int idx; } orc_mt19937;
void orc_mt_seed(orc_mt19937* g, uint32_t seed);
uint32_t orc_mt_next(orc_mt19937* g);

/* ---- SampleConsensusModel sampling (sac_model.h: getSamples / drawIndexSample / rnd) ----- */
typedef struct { orc_mt19937 rng; int32_t* shuffled; size_t n; } orc_sampler;
int orc_sampler_init(orc_sampler* s, size_t n, uint32_t seed);
void orc_sampler_free(orc_sampler* s);
void orc_sampler_draw(orc_sampler* s, int32_t idx[3]);
void orc_sampler_draw_k(orc_sampler* s, int sample_size, int32_t* idx);
/* Convenience for tests: the first n_draws raw draws for a cloud of n points. */
int orc_draw_sequence(size_t n, uint32_t seed, int n_draws, int32_t* triples /* 3*n_draws */);

/* ---- SampleConsensusModelPlane (sac_model_plane.hpp) ------------------------------------- */
int orc_is_sample_good(const orc_point* cloud, const int32_t idx[3]);
int orc_compute_model(const orc_point* cloud, const int32_t idx[3], float coeff[4]);
float orc_signed_distance(const float coeff[4], const orc_point* p, int dot_order);
void orc_residuals(const orc_point* cloud, size_t n, const float coeff[4], int dot_order, float* out);
int64_t orc_count_within(const orc_point* cloud, size_t n, const float coeff[4], double t, int dot_order);
/* Same loop, OpenMP-parallel over points (NOT how PCL 1.8 runs; used by the multi-thread CPU arm). */
int64_t orc_count_within_mt(const orc_point* cloud, size_t n, const float coeff[4], double t, int dot_order);
size_t orc_select_within(const orc_point* cloud, size_t n, const float coeff[4], double t, int dot_order,
                         int32_t* out);
/* counts[k] for K models, scalar PCL loop per model; threads > 1 uses orc_count_within_mt. */
void orc_count_batch(const orc_point* cloud, size_t n, const float* coeffs /* 4*K */, int K, double t,
                     int dot_order, int threads, int32_t* counts);

/* ---- optimizeModelCoefficients ------------------------------------------------------------ */
int orc_refit_pcl_float(const orc_point* cloud, const int32_t* idx, size_t n_idx, const float coeff_in[4],
                        float coeff_out[4]);
/* Extent exponent of a cloud: s such that |x - pivot| * 2^s < 2^30 for every finite point. */
int orc_fixed_scale_exp(const orc_point* cloud, size_t n);
/* moments_out (optional, 19 x int64): n, Sx,Sy,Sz, then (hi,lo-as-int64) pairs... see .c */
int orc_refit_fixed(const orc_point* cloud, const int32_t* idx, size_t n_idx, const float pivot[3],
                    int scale_exp, const float coeff_in[4], float coeff_out[4], int64_t* moments_out);
/* The host half of the fixed refit alone: integer moments -> plane.  Exposed so tests can feed it
 * the device's moments.  m = {n, Sx, Sy, Sz, Sxx_hi, Sxx_lo, Sxy_hi, Sxy_lo, ... Szz_hi, Szz_lo}
 * with S_ab = hi * 2^32 + lo. */
int orc_plane_from_moments(const int64_t m[16], const float pivot[3], int scale_exp, float coeff_out[4]);

/* ---- staging steps of the reference's preProcess (Dialog/PlaneDetect.h:449-481) --------------
 * pcl::removeNaNFromPointCloud: finite points in order; map[i] = source index.  Returns the count. */
size_t orc_remove_nonfinite(const orc_point* cloud, size_t n, orc_point* out, int32_t* map);
/* Exactly rounded centroid of the finite points (integer sums about the bbox corner on the refit grid —
 * the order-independent form of preProcess's sequential float sum).  Returns 0 when there is no finite point. */
int orc_centroid_exact(const orc_point* cloud, size_t n, float centroid[3]);
/* points[i] -= centroid, one float subtraction per coordinate. */
void orc_translate(orc_point* cloud, size_t n, const float centroid[3]);

/* projPoint2Plane (Dialog/PlaneDetect.h:1442-1448) applied to cloud[idx[i]]: lambda = 2.0 * (a*x + b*y + c*z + d)
 * stored as float, dest = src - lambda / 2.0 * n evaluated in double and stored as float. */
void orc_project_points(const orc_point* cloud, const int32_t* idx, size_t n_idx, const float coeff[4], orc_point* out);

/* ---- postProcessPlanes re-absorption (Dialog/PlaneDetect.h:1454-1580; REFERENCE code, not PCL) --------------
 * The reference's own use of the threshold test + peel: every still-unclaimed point is tested against every
 * newly found plane polygon with isPointInPoly (:1891-1955) and claimed by each plane that contains it
 * (no break: a point can join several planes); the unclaimed rest becomes the new source_cloud (:1560-1566).
 *
 * Arithmetic restated op for op in FP32 without contraction (MSVC v140 x64 /fp:precise):
 *   getInfoBetPointAndPlane / projPoint2Plane / distP2P (:1442-1448, 2019-2023, 203-207); dist > T rejects
 *   (non-strict: dist == T stays a candidate).
 *   isBothLineSegsIntersect (:1957-2016).
 * CHOICES (libraries the reference links but does not vendor; flagged for re-verification):
 *   Eigen Vector3f dot / squaredNorm reduce as e0 + (e1 + e2) (redux_novec_unroller halves the range);
 *   normalize() divides each component by sqrt(squaredNorm) when squaredNorm > 0 (Eigen >= 3.3);
 *   pow(a, 0.5f) is taken as the correctly rounded sqrtf(a);
 *   rand() is the MSVC CRT generator (holdrand = holdrand * 214013 + 2531011; (holdrand >> 16) & 0x7fff),
 *   re-seeded by srand(seed) at every isPointInPoly call as the reference does with srand(time(0)), so the ten
 *   ray edges of a plane are a function of (seed, border size) alone. */
void orc_msvc_rand_edges(unsigned seed, int border_size, int edges[10]);
int orc_segs_intersect(const orc_point* pa, const orc_point* pb, const orc_point* pc, const orc_point* pd);
/* 1 if isPointInPoly(p, plane{coeff, border}) with T_dist_point_plane = t, srand(seed). */
int orc_point_in_poly(const orc_point* p, const float coeff[4], const orc_point* border, int n_border, float t,
                      unsigned seed);
/* The loop of postProcessPlanes over an unclaimed cloud: planes j in [0, n_planes) with coeffs[4j..] and border
 * vertices border[border_offsets[j] .. border_offsets[j+1]).  absorbed: per plane the indices of the points it
 * claimed, ascending, concatenated; plane_offsets: n_planes + 1; remaining_idx: indices of the unclaimed points,
 * ascending.  Returns 0, or -1 when a capacity is too small or a border is empty. */
int orc_reabsorb(const orc_point* cloud, size_t n, const float* coeffs, const orc_point* border,
                 const size_t* border_offsets, int n_planes, float t, unsigned seed, int32_t* absorbed, size_t absorbed_cap,
                 size_t* plane_offsets, int32_t* remaining_idx, size_t* n_remaining);

/* ---- pcl::NormalEstimationOMP with a radius search (Dialog/PlaneDetect.h:515-545), see pr_oracle.c --------------
 * out: 4 floats per point (normal_x, normal_y, normal_z, curvature), NaN where PCL yields NaN; n_neighbors (optional):
 * neighbours found per point (the point itself included).  Brute force over all pairs: small clouds only. */
enum { ORC_NORMALS_PCL_FLOAT = 0, ORC_NORMALS_FIXED = 1 };
int orc_normals_scale_exp(double radius);
int orc_estimate_normals(const orc_point* cloud, size_t n, double radius, const float vp[3], int mode, float* out,
                         int32_t* n_neighbors);

/* ---- clusterFilt (Dialog/PlaneDetect.h:1582-1656): keep[i] = 0 for the points of radius-graph components with at most
 * max_small_cluster points.  Brute force over all pairs: small clouds only. */
int orc_cluster_filter(const orc_point* cloud, size_t n, double radius, int max_small_cluster, uint8_t* keep);

/* ---- RandomSampleConsensus::computeModel + SACSegmentation::segment ----------------------- */
int orc_segment(const orc_point* cloud, size_t n, const orc_params* prm, int scale_exp_or_min,
                float coeff[4], int32_t* inliers /* cap n */, size_t* n_inliers, orc_trace* trace);

/* ---- segment + ExtractIndices peel loop ---------------------------------------------------
 * coeffs: 4*max_planes; inlier_cur: indices into the cloud of that round (what PCL's loop yields);
 * inlier_orig: the same points as indices into the input cloud; plane_offsets: max_planes+1;
 * remaining (optional): cap n points; traces (optional): max_planes+1 entries (the last, rejected,
 * segment call is traced too). */
int orc_extract_planes(const orc_point* cloud, size_t n, const orc_params* prm, float* coeffs,
                       int32_t* inlier_cur, int32_t* inlier_orig, size_t idx_cap, size_t* plane_offsets,
                       int* n_planes, orc_point* remaining, size_t* n_remaining, orc_trace* traces);

/* ---- SACMODEL_LINE exactly as the reference calls it (Dialog/SimplifyVerticesSize.cpp:64-67): same sampler, same
 * computeModel loop as the plane model; coeff = (point, direction). */
int orc_segment_line(const orc_point* cloud, size_t n, const orc_params* prm, float coeff[6], int32_t* inliers, size_t* n_inliers,
                     orc_trace* trace);
/* ---- clusterFilt restated literally from Dialog/PlaneDetect.h:1598-1634 (BFS, first radius-search hit skipped), to
 * enumerate where it differs from the connected-components definition. */
int orc_cluster_filter_reference_bfs(const orc_point* cloud, size_t n, double radius, int max_small_cluster, int ties_by_index,
                                     uint8_t* keep);

#ifdef __cplusplus
}
#endif
#endif
