"""dialog_b200 — B200 (sm_100a) backend for czh55/Dialog's plane-detection path (RANSAC + peel)."""
from .plane_detect import (DOT_FMA, DOT_PCL_SSE2, REFIT_FIXED, REFIT_PCL_FLOAT, SCORER_BRUTE, SCORER_HIER, Extraction, PinnedArray, Plane, PlaneRansac,  # noqa: F401
                           PlaneRansacError, LOOP_STAGE_NAMES,
                           detect_planes, host_draw_triples, host_draw_triples_parallel, host_plane_from_moments, host_rand_edges, host_replay,
                           host_shard_range, make_params)
