"""ctypes binding of libplane_ransac.so (include/plane_ransac.h).

The library is the product: if it is missing this module raises — there is no Python or CPU
implementation of the kernels to fall back to.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libplane_ransac.so")

UNIQUE_ID_BYTES = 128
DOT_PCL_SSE2 = 0
DOT_FMA = 1
SCORER_BRUTE = 0
SCORER_HIER = 1
REFIT_FIXED = 0
REFIT_PCL_FLOAT = 1
STAGE_REMOVE_NONFINITE = 1
STAGE_TRANSLATE_CENTROID = 2

# every symbol include/plane_ransac.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "plane_ransac_abi_version", "plane_ransac_last_error", "plane_ransac_default_params",
    "plane_ransac_create", "plane_ransac_destroy", "plane_ransac_set_cloud", "plane_ransac_set_cloud_device",
    "plane_ransac_set_cloud_ex", "plane_ransac_set_cloud_async", "plane_ransac_staged_source_indices",
    "plane_ransac_cloud_size", "plane_ransac_score", "plane_ransac_segment_one", "plane_ransac_extract_planes",
    "plane_ransac_plane_points", "plane_ransac_set_round_loop", "plane_ransac_remaining", "plane_ransac_reabsorb", "plane_ransac_estimate_normals", "plane_ransac_cluster_filter", "plane_ransac_restage_remaining", "plane_ransac_set_cloud_batch", "plane_ransac_segment_batch", "plane_ransac_segment_batch_lists",
    "plane_ransac_comm_unique_id", "plane_ransac_comm_init", "plane_ransac_comm_p2p_enabled", "plane_ransac_shard_info",
    "plane_ransac_load_pcd", "plane_ransac_host_alloc", "plane_ransac_host_free", "plane_ransac_host_register", "plane_ransac_host_unregister", "plane_ransac_profile_enable", "plane_ransac_profile_reset", "plane_ransac_profile_get", "plane_ransac_round_timeline",
    "plane_ransac_timer_start", "plane_ransac_timer_stop", "plane_ransac_measure_ffma_peak", "plane_ransac_measure_copy_bw", "plane_ransac_flush_l2",
    "plane_ransac_host_draw_triples", "plane_ransac_host_draw_triples_parallel", "plane_ransac_host_replay", "plane_ransac_host_shard_range",
    "plane_ransac_host_plane_from_moments", "plane_ransac_host_plane_from_pcl_float_sums", "plane_ransac_host_rand_edges",
]


class PrParams(C.Structure):
    _fields_ = [
        ("distance_threshold", C.c_double),
        ("max_iterations", C.c_int),
        ("min_plane_size", C.c_int),
        ("probability", C.c_double),
        ("optimize_coefficients", C.c_int),
        ("seed", C.c_uint),
        ("max_planes", C.c_int),
        ("dot_order", C.c_int),
        ("scorer", C.c_int),
        ("refit_mode", C.c_int),
    ]


class PrSegmentInfo(C.Structure):
    _fields_ = [
        ("ok", C.c_int),
        ("iterations", C.c_int),
        ("draws", C.c_int),
        ("skipped", C.c_int),
        ("best_sample", C.c_int * 3),
        ("best_count", C.c_int),
        ("raw_coeff", C.c_float * 4),
        ("n_inliers_raw", C.c_int),
        ("n_inliers", C.c_int),
        ("scale_exp", C.c_int),
        ("n_scored", C.c_int),
        ("n_cloud", C.c_longlong),
    ]


LOOP_STAGES = 9  # PR_LOOP_STAGES
LOOP_STAGE_NAMES = ("draw", "draw_resolve", "models", "score", "decide", "refit", "finish", "peel", "advance")


class PrProfile(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ms_stage", "ms_models", "ms_score", "ms_refit", "ms_compact", "ms_other")] + \
               [(n, C.c_longlong) for n in ("launches_stage", "launches_models", "launches_score", "launches_refit",
                                            "launches_compact", "launches_other", "pairs_scored", "points_refit",
                                            "points_compact", "bytes_compact", "bytes_refit")] + \
               [(n, C.c_double) for n in ("host_ms_sampling", "host_ms_replay", "host_ms_wait", "host_ms_total")] + \
               [(n, C.c_longlong) for n in ("points_kept", "points_peeled")] + \
               [("p2p_wait_ms", C.c_double * 4), ("p2p_exchanges", C.c_longlong * 4)] + \
               [("loop_ms", C.c_double * LOOP_STAGES), ("loop_rounds", C.c_longlong)]


class PlaneRansacError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"plane_ransac error {code}: {message}")
        self.code = code


_lib = None


def load():
    """Load the shared library, building nothing: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dialog_b200.build` "
            "(there is no CPU or PyTorch fallback for the CUDA kernels)")
    L = C.CDLL(LIB_PATH)
    vp, sz = C.c_void_p, C.c_size_t
    L.plane_ransac_abi_version.restype = C.c_int
    L.plane_ransac_last_error.restype = C.c_char_p
    L.plane_ransac_default_params.argtypes = [C.POINTER(PrParams)]
    L.plane_ransac_default_params.restype = None
    L.plane_ransac_create.argtypes = [C.POINTER(vp), C.c_int]
    L.plane_ransac_destroy.argtypes = [vp]
    L.plane_ransac_destroy.restype = None
    L.plane_ransac_set_cloud.argtypes = [vp, vp, sz]
    L.plane_ransac_set_cloud_device.argtypes = [vp, vp, sz]
    L.plane_ransac_set_cloud_async.argtypes = [vp, vp, sz]
    L.plane_ransac_set_cloud_ex.argtypes = [vp, vp, sz, C.c_uint, C.POINTER(sz), vp]
    L.plane_ransac_staged_source_indices.argtypes = [vp, vp, sz]
    L.plane_ransac_cloud_size.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    L.plane_ransac_score.argtypes = [vp, vp, C.c_int, C.c_double, C.c_int, vp, vp, vp]
    L.plane_ransac_segment_one.argtypes = [vp, C.POINTER(PrParams), vp, vp, sz, C.POINTER(sz), C.POINTER(PrSegmentInfo)]
    L.plane_ransac_extract_planes.argtypes = [vp, C.POINTER(PrParams), vp, vp, vp, sz, vp, C.POINTER(C.c_int), vp]
    L.plane_ransac_plane_points.argtypes = [vp, C.c_int, C.c_int, vp, sz, C.POINTER(sz)]
    L.plane_ransac_remaining.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.plane_ransac_reabsorb.argtypes = [vp, vp, vp, vp, C.c_int, C.c_float, C.c_uint, vp, vp, sz, vp, C.POINTER(sz)]
    L.plane_ransac_restage_remaining.argtypes = [vp]
    L.plane_ransac_estimate_normals.argtypes = [vp, C.c_double, vp, vp, sz, vp]
    L.plane_ransac_cluster_filter.argtypes = [vp, C.c_double, C.c_int, C.POINTER(sz), C.POINTER(sz)]
    L.plane_ransac_set_cloud_batch.argtypes = [vp, vp, sz, sz]
    L.plane_ransac_segment_batch.argtypes = [vp, C.POINTER(PrParams), vp, vp, vp]
    L.plane_ransac_comm_unique_id.argtypes = [vp]
    L.plane_ransac_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.plane_ransac_comm_p2p_enabled.argtypes = [vp]
    L.plane_ransac_shard_info.argtypes = [vp] + [C.POINTER(C.c_longlong)] * 4
    L.plane_ransac_host_alloc.argtypes = [sz, C.POINTER(vp)]
    L.plane_ransac_host_free.argtypes = [vp]
    L.plane_ransac_load_pcd.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(sz)]
    L.plane_ransac_host_register.argtypes = [vp, sz]
    L.plane_ransac_host_unregister.argtypes = [vp]
    L.plane_ransac_profile_enable.argtypes = [vp, C.c_int]
    L.plane_ransac_profile_reset.argtypes = [vp]
    L.plane_ransac_profile_get.argtypes = [vp, C.POINTER(PrProfile)]
    L.plane_ransac_round_timeline.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.plane_ransac_timer_start.argtypes = [vp]
    L.plane_ransac_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
    L.plane_ransac_measure_ffma_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.plane_ransac_measure_copy_bw.argtypes = [vp, sz, C.POINTER(C.c_double)]
    L.plane_ransac_flush_l2.argtypes = [vp]
    L.plane_ransac_host_draw_triples.argtypes = [sz, C.c_uint, C.c_int, vp]
    L.plane_ransac_segment_batch_lists.argtypes = [vp, C.POINTER(PrParams), vp, vp, vp, sz, vp, vp]
    L.plane_ransac_set_round_loop.argtypes = [vp, C.c_int]
    L.plane_ransac_host_draw_triples_parallel.argtypes = [sz, C.c_uint, C.c_int, vp, C.POINTER(C.c_int)]
    L.plane_ransac_host_replay.argtypes = [vp, vp, C.c_int, C.c_longlong, C.c_int, C.c_double] + [C.POINTER(C.c_int)] * 5
    L.plane_ransac_host_shard_range.argtypes = [C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.plane_ransac_host_plane_from_moments.argtypes = [vp, vp, C.c_int, vp]
    L.plane_ransac_host_plane_from_pcl_float_sums.argtypes = [vp, C.c_longlong, vp]
    L.plane_ransac_host_rand_edges.argtypes = [C.c_uint, C.c_int, vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("plane_ransac_abi_version",):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise PlaneRansacError(rc, load().plane_ransac_last_error().decode("utf-8", "replace"))
