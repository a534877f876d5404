"""Host-side mirror of the PlaneDetect call surface for the RANSAC backend.

The reference drives plane detection as a stage between estimateNormal() and polyPlanes()
(Dialog/PCLViewer.cpp:1183-1226): input is the global PointCloudT::Ptr source_cloud
(Dialog/PlaneDetect.h:104), output is one Plane{coeff, points_set} per plane appended to
plane_clouds_final (Dialog/HeaderFile.h:81-98) and source_cloud rebuilt from the unclaimed points
(Dialog/PlaneDetect.h:1560-1566).  `detect_planes` has that shape; `PlaneRansac` is the thin object
over the C ABI (include/plane_ransac.h) that the tests and the bench harness drive.  The C++ shim with
the same surface is include/PlaneDetectRansac.h.

All device work happens in libplane_ransac.so; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import (DOT_FMA, DOT_PCL_SSE2, REFIT_FIXED, REFIT_PCL_FLOAT, SCORER_BRUTE, SCORER_HIER, PlaneRansacError, PrParams, PrProfile,  # noqa: F401
                   PrSegmentInfo, LOOP_STAGE_NAMES)


def make_params(distance_threshold: float = 0.1, max_iterations: int = 50, min_plane_size: int = 500,
                probability: float = 0.99, optimize_coefficients: bool = True, seed: int = 12345,
                max_planes: int = 64, dot_order: int = DOT_FMA, scorer: int = 0, refit_mode: int = 0) -> PrParams:
    """Defaults: T_dist_point_plane / T_num_of_single_plane from Dialog/config.txt:29,20 and PCL's
    SACSegmentation defaults for the knobs config.txt has no key for."""
    return PrParams(float(distance_threshold), int(max_iterations), int(min_plane_size), float(probability),
                    int(bool(optimize_coefficients)), int(seed), int(max_planes), int(dot_order), int(scorer),
                    int(refit_mode))


def as_cloud(points: np.ndarray) -> np.ndarray:
    """(N,4) float32 C-contiguous rows laid out like pcl::PointXYZ (x, y, z, 1)."""
    a = np.asarray(points)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError("cloud must have shape (N,3) or (N,4)")
    if a.shape[1] == 3:
        b = np.ones((a.shape[0], 4), np.float32)
        b[:, :3] = a
        return b
    return np.ascontiguousarray(a, dtype=np.float32)


class PinnedArray:
    """numpy view of page-locked host memory owned by the library (plane_ransac_host_alloc)."""

    def __init__(self, shape, dtype):
        self._L = _lib.load()
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        _lib.check(self._L.plane_ransac_host_alloc(max(nbytes, 1), C.byref(p)))
        self._p = p
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)

    @property
    def ptr(self) -> int:
        return self._p.value

    def free(self):
        if self._p is not None:
            self.array = None
            self._L.plane_ransac_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class Plane:
    """== struct Plane (Dialog/HeaderFile.h:81-88) restricted to what this path fills."""
    coeff: np.ndarray                 # (4,) a, b, c, d
    inliers_cur: np.ndarray           # indices into the cloud of the round that found it
    inliers_orig: np.ndarray          # indices into the staged cloud (== points_set)
    info: PrSegmentInfo = None


@dataclass
class Extraction:
    planes: list = field(default_factory=list)
    infos: list = field(default_factory=list)   # one per segment call, the final rejected one included

    @property
    def coeffs(self) -> np.ndarray:
        return np.array([p.coeff for p in self.planes], np.float32).reshape(-1, 4)


class PlaneRansac:
    """One CUDA device + stream behind the C ABI.  Not thread-safe (like the reference's GUI-thread path)."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.plane_ransac_create(C.byref(h), device))
        self._h = h
        self._idx_buf = None
        self._batch = 0
        self._batch_n = 0

    def close(self):
        if getattr(self, "_idx_buf", None):
            for b in self._idx_buf:
                b.free()
            self._idx_buf = None
        if getattr(self, "_h", None):
            self._L.plane_ransac_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- staging ----
    def set_cloud(self, points: np.ndarray) -> int:
        a = as_cloud(points)
        _lib.check(self._L.plane_ransac_set_cloud(self._h, a.ctypes.data_as(C.c_void_p), a.shape[0]))
        return a.shape[0]

    def set_cloud_preprocessed(self, points: np.ndarray, remove_nonfinite: bool = True, translate: bool = True):
        """Stage with preProcess's NaN removal / centroid translation fused in (Dialog/PlaneDetect.h:449-481).
        Returns (points kept, centroid, source index of every staged point)."""
        a = as_cloud(points)
        flags = (_lib.STAGE_REMOVE_NONFINITE if remove_nonfinite else 0) | (_lib.STAGE_TRANSLATE_CENTROID if translate else 0)
        kept = C.c_size_t(0)
        cen = np.zeros(3, np.float32)
        _lib.check(self._L.plane_ransac_set_cloud_ex(self._h, a.ctypes.data_as(C.c_void_p), a.shape[0], flags,
                                                     C.byref(kept), cen.ctypes.data_as(C.c_void_p)))
        src = np.empty(max(kept.value, 1), np.int32)
        _lib.check(self._L.plane_ransac_staged_source_indices(self._h, src.ctypes.data_as(C.c_void_p), src.size))
        return kept.value, cen, src[: kept.value]

    def staged_source_indices(self) -> np.ndarray:
        """Index in the caller's array of every staged point (identity unless staged with a filter or restaged)."""
        n, _ = self.cloud_size()
        src = np.empty(max(n, 1), np.int32)
        _lib.check(self._L.plane_ransac_staged_source_indices(self._h, src.ctypes.data_as(C.c_void_p), src.size))
        return src[:n]

    def set_cloud_ptr(self, host_ptr: int, n: int, overlap: bool = False) -> None:
        """Stage from a raw host address (e.g. a pinned torch tensor's data_ptr()).  overlap=True queues the upload
        and lets the next extract / segment call score the chunks as they land (plane_ransac_set_cloud_async: the
        buffer must be pinned and stay untouched until that call returns)."""
        f = self._L.plane_ransac_set_cloud_async if overlap else self._L.plane_ransac_set_cloud
        _lib.check(f(self._h, C.c_void_p(host_ptr), n))

    def set_cloud_device_ptr(self, dev_ptr: int, n: int) -> None:
        _lib.check(self._L.plane_ransac_set_cloud_device(self._h, C.c_void_p(dev_ptr), n))

    def cloud_size(self):
        a, b = C.c_size_t(0), C.c_size_t(0)
        _lib.check(self._L.plane_ransac_cloud_size(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- scoring hook ----
    def score(self, triples: np.ndarray, distance_threshold: float, dot_order: int = DOT_FMA, want_models: bool = False):
        t = np.ascontiguousarray(triples, np.int32).reshape(-1, 3)
        K = t.shape[0]
        counts = np.zeros(K, np.int32)
        coeffs = np.zeros((K, 4), np.float32) if want_models else None
        good = np.zeros(K, np.uint8) if want_models else None
        _lib.check(self._L.plane_ransac_score(
            self._h, t.ctypes.data_as(C.c_void_p), K, float(distance_threshold), dot_order,
            counts.ctypes.data_as(C.c_void_p),
            coeffs.ctypes.data_as(C.c_void_p) if want_models else None,
            good.ctypes.data_as(C.c_void_p) if want_models else None))
        return (counts, coeffs, good.astype(bool)) if want_models else counts

    # ---- segment / extract ----
    def segment_one(self, params: PrParams):
        n_staged, _ = self.cloud_size()
        coeff = np.zeros(4, np.float32)
        inl = np.empty(max(n_staged, 1), np.int32)
        n = C.c_size_t(0)
        info = PrSegmentInfo()
        _lib.check(self._L.plane_ransac_segment_one(self._h, C.byref(params), coeff.ctypes.data_as(C.c_void_p),
                                                    inl.ctypes.data_as(C.c_void_p), inl.size, C.byref(n), C.byref(info)))
        return coeff, inl[: n.value].copy(), info

    def extract_planes(self, params: PrParams, want_indices: bool = True, copy: bool = True) -> Extraction:
        """copy=False returns views into the context's pinned result buffers (valid until the next call)."""
        n_staged, _ = self.cloud_size()
        mp = max(int(params.max_planes), 0)   # a negative value is rejected by the library
        coeffs = np.zeros((max(mp, 1), 4), np.float32)
        offs = np.zeros(mp + 1, np.uintp)
        npl = C.c_int(0)
        infos = (PrSegmentInfo * (mp + 1))()
        cur = orig = None
        if want_indices:
            # page-locked result buffers, kept across calls (they are overwritten by the next call)
            if self._idx_buf is None or self._idx_buf[0].array.size < max(n_staged, 1):
                self._idx_buf = (PinnedArray((max(n_staged, 1),), np.int32), PinnedArray((max(n_staged, 1),), np.int32))
            cur, orig = self._idx_buf[0].array, self._idx_buf[1].array
        _lib.check(self._L.plane_ransac_extract_planes(
            self._h, C.byref(params), coeffs.ctypes.data_as(C.c_void_p),
            cur.ctypes.data_as(C.c_void_p) if want_indices else None,
            orig.ctypes.data_as(C.c_void_p) if want_indices else None,
            n_staged if want_indices else 0, offs.ctypes.data_as(C.c_void_p), C.byref(npl), infos))
        P = npl.value
        o = [int(v) for v in offs[: P + 1]]
        ex = Extraction()
        for k in range(P):
            ex.planes.append(Plane(coeffs[k].copy(),
                                   (cur[o[k]: o[k + 1]].copy() if copy else cur[o[k]: o[k + 1]]) if want_indices else None,
                                   (orig[o[k]: o[k + 1]].copy() if copy else orig[o[k]: o[k + 1]]) if want_indices else None,
                                   infos[k]))
        ex.infos = [infos[k] for k in range(min(P + 1, mp))]
        return ex

    # ---- batch of equal-sized small clouds (BASELINE config 5) ----
    def set_cloud_batch(self, clouds: np.ndarray) -> None:
        """clouds: (n_clouds, n_per_cloud, 3|4) float32."""
        a = np.asarray(clouds)
        if a.ndim != 3 or a.shape[2] not in (3, 4):
            raise ValueError("batch must have shape (n_clouds, n_per_cloud, 3|4)")
        flat = as_cloud(a.reshape(-1, a.shape[2]))
        _lib.check(self._L.plane_ransac_set_cloud_batch(self._h, flat.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1]))
        self._batch = a.shape[0]
        self._batch_n = a.shape[1]

    def set_cloud_batch_ptr(self, host_ptr: int, n_clouds: int, n_per_cloud: int) -> None:
        _lib.check(self._L.plane_ransac_set_cloud_batch(self._h, C.c_void_p(host_ptr), n_clouds, n_per_cloud))
        self._batch = n_clouds
        self._batch_n = n_per_cloud

    def segment_batch(self, params: PrParams, want_infos: bool = True, want_lists: bool = False, lists_buf: np.ndarray = None):
        """One segment() per cloud: (coeffs (n_clouds,4), n_inliers (n_clouds,), infos) and, with want_lists, a fourth
        item: the list of per-cloud ascending inlier index arrays (views into one buffer).  lists_buf: the caller's int32
        buffer for the lists (a PinnedArray's .array makes the read-back a single DMA; too small: PlaneRansacError -4)."""
        nc = self._batch
        coeffs = np.zeros((nc, 4), np.float32)
        cnt = np.zeros(nc, np.int32)
        infos = (PrSegmentInfo * nc)() if want_infos else None
        if not want_lists:
            _lib.check(self._L.plane_ransac_segment_batch(self._h, C.byref(params), coeffs.ctypes.data_as(C.c_void_p),
                                                          cnt.ctypes.data_as(C.c_void_p), infos))
            return coeffs, cnt, infos
        offs = np.zeros(nc + 1, np.uintp)
        if lists_buf is not None:
            buf = lists_buf.reshape(-1)
            assert buf.dtype == np.int32 and buf.flags["C_CONTIGUOUS"]
            _lib.check(self._L.plane_ransac_segment_batch_lists(self._h, C.byref(params), coeffs.ctypes.data_as(C.c_void_p),
                                                                cnt.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), buf.size,
                                                                offs.ctypes.data_as(C.c_void_p), infos))
            return coeffs, cnt, infos, [buf[int(offs[k]): int(offs[k + 1])] for k in range(nc)]
        cap = max(1, self._batch_n * nc // 2)
        for _ in range(2):
            buf = np.empty(cap, np.int32)
            rc = self._L.plane_ransac_segment_batch_lists(self._h, C.byref(params), coeffs.ctypes.data_as(C.c_void_p),
                                                          cnt.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), cap,
                                                          offs.ctypes.data_as(C.c_void_p), infos)
            if rc != -4:
                break
            cap = int(offs[nc])    # PR_ERR_CAPACITY: the offsets say how much is needed
        _lib.check(rc)
        return coeffs, cnt, infos, [buf[int(offs[k]): int(offs[k + 1])] for k in range(nc)]

    def plane_points(self, k: int, project: bool = False) -> np.ndarray:
        """Plane::points_set of plane k of the last extract call; project=True gives the cloud polyPointCloud
        hands to pcl::ConcaveHull (points projected onto the plane, Dialog/PlaneDetect.h:1391-1397)."""
        n = C.c_size_t(0)
        _lib.check(self._L.plane_ransac_plane_points(self._h, k, int(project), None, 0, C.byref(n)))
        out = np.empty((max(n.value, 1), 4), np.float32)
        _lib.check(self._L.plane_ransac_plane_points(self._h, k, int(project), out.ctypes.data_as(C.c_void_p),
                                                     out.shape[0], C.byref(n)))
        return out[: n.value]

    def remaining(self) -> np.ndarray:
        _, n_cur = self.cloud_size()
        out = np.empty((max(n_cur, 1), 4), np.float32)
        n = C.c_size_t(0)
        _lib.check(self._L.plane_ransac_remaining(self._h, out.ctypes.data_as(C.c_void_p), out.shape[0], C.byref(n)))
        return out[: n.value]

    def estimate_normals(self, radius: float, viewpoint=(0.0, 0.0, 0.0), want_counts: bool = False):
        """estimateNormal() of the reference (pcl::NormalEstimationOMP, radius search) on the current cloud:
        (n,4) float32 rows (normal_x, normal_y, normal_z, curvature), NaN where PCL yields NaN; with want_counts also the
        number of neighbours per point."""
        _, n = self.cloud_size()
        out = np.empty((max(n, 1), 4), np.float32)
        cnt = np.zeros(max(n, 1), np.int32) if want_counts else None
        vp = np.ascontiguousarray(viewpoint, np.float32)
        _lib.check(self._L.plane_ransac_estimate_normals(self._h, float(radius), vp.ctypes.data_as(C.c_void_p),
                                                         out.ctypes.data_as(C.c_void_p), out.shape[0],
                                                         cnt.ctypes.data_as(C.c_void_p) if want_counts else None))
        return (out[:n], cnt[:n]) if want_counts else out[:n]

    def cluster_filter(self, radius: float, max_small_cluster: int):
        """clusterFilt() (Dialog/PlaneDetect.h:1582-1656): drop every connected component of the radius graph with at
        most max_small_cluster points from the current cloud.  Returns (points removed, points left)."""
        a, b = C.c_size_t(0), C.c_size_t(0)
        _lib.check(self._L.plane_ransac_cluster_filter(self._h, float(radius), int(max_small_cluster), C.byref(a), C.byref(b)))
        return a.value, b.value

    def restage_remaining(self) -> None:
        """The cloud left by the last extract / reabsorb call becomes the staged cloud (the reference's "run again",
        Dialog/PCLViewer.cpp:1120-1178); staged_source_indices() maps it to the caller's array."""
        _lib.check(self._L.plane_ransac_restage_remaining(self._h))

    def reabsorb(self, coeffs, borders, distance_threshold: float = 0.1, rand_seed: int = 0):
        """postProcessPlanes' re-absorption pass (Dialog/PlaneDetect.h:1530-1566) over the current cloud: every point is
        tested against every plane polygon with the reference's isPointInPoly; claimed points leave the cloud.
        coeffs: (P,4); borders: P arrays (nb_j, 3|4) of polygon vertices (Plane::border); rand_seed: what the reference
        passes to srand (time(0)).  Returns (per-plane ascending index lists into the cloud before the call, the same as
        indices into the staged cloud, number of points left)."""
        co = np.ascontiguousarray(coeffs, np.float32).reshape(-1, 4)
        P = co.shape[0]
        if len(borders) != P:
            raise ValueError("one border polygon per plane")
        bd = np.concatenate([as_cloud(b) for b in borders]) if P else np.zeros((0, 4), np.float32)
        offs = np.zeros(P + 1, np.uintp)
        if P:
            offs[1:] = np.cumsum([len(b) for b in borders])
        _, n_cur = self.cloud_size()
        cap = max(1, n_cur)
        po = np.zeros(P + 1, np.uintp)
        nrem = C.c_size_t(0)
        vp = C.c_void_p
        while True:
            cur = np.empty(cap, np.int32)
            orig = np.empty(cap, np.int32)
            try:
                _lib.check(self._L.plane_ransac_reabsorb(self._h, co.ctypes.data_as(vp), bd.ctypes.data_as(vp), offs.ctypes.data_as(vp), P,
                                                         float(distance_threshold), int(rand_seed) & 0xFFFFFFFF, cur.ctypes.data_as(vp),
                                                         orig.ctypes.data_as(vp), cap, po.ctypes.data_as(vp), C.byref(nrem)))
                break
            except PlaneRansacError as e:  # a point can join several planes: the lists can outgrow the cloud
                if e.code != -4 or cap >= max(1, n_cur) * max(P, 1):
                    raise
                cap = max(1, n_cur) * max(P, 1)
        o = [int(v) for v in po]
        return ([cur[o[j]: o[j + 1]].copy() for j in range(P)], [orig[o[j]: o[j + 1]].copy() for j in range(P)], int(nrem.value))

    # ---- sharding ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        _lib.check(_lib.load().plane_ransac_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes) -> None:
        buf = C.create_string_buffer(unique_id, _lib.UNIQUE_ID_BYTES)
        _lib.check(self._L.plane_ransac_comm_init(self._h, n_ranks, rank, buf))

    def p2p_enabled(self) -> bool:
        """True when the per-round exchanges run as peer-memory kernels over NVLink instead of NCCL collectives."""
        return bool(self._L.plane_ransac_comm_p2p_enabled(self._h))

    def shard_info(self):
        v = [C.c_longlong(0) for _ in range(4)]
        _lib.check(self._L.plane_ransac_shard_info(self._h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    # ---- measurement ----
    def profile_enable(self, on: bool = True):
        _lib.check(self._L.plane_ransac_profile_enable(self._h, int(on)))

    def profile_reset(self):
        _lib.check(self._L.plane_ransac_profile_reset(self._h))

    def profile(self) -> PrProfile:
        p = PrProfile()
        _lib.check(self._L.plane_ransac_profile_get(self._h, C.byref(p)))
        return p

    def round_timeline(self) -> np.ndarray:
        """(rounds, LOOP_STAGES + 1) uint64 %globaltimer stamps of the last extract_planes call's device-loop rounds."""
        n = C.c_size_t(0)
        _lib.check(self._L.plane_ransac_round_timeline(self._h, None, 0, C.byref(n)))
        out = np.zeros((n.value, _lib.LOOP_STAGES + 1), dtype=np.uint64)
        if n.value:
            _lib.check(self._L.plane_ransac_round_timeline(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def timer_start(self):
        _lib.check(self._L.plane_ransac_timer_start(self._h))

    def timer_stop(self) -> float:
        v = C.c_double(0)
        _lib.check(self._L.plane_ransac_timer_stop(self._h, C.byref(v)))
        return v.value

    def measure_ffma_peak(self) -> float:
        v = C.c_double(0)
        _lib.check(self._L.plane_ransac_measure_ffma_peak(self._h, C.byref(v)))
        return v.value

    def measure_copy_bw(self, nbytes: int = 1 << 30) -> float:
        v = C.c_double(0)
        _lib.check(self._L.plane_ransac_measure_copy_bw(self._h, nbytes, C.byref(v)))
        return v.value

    def set_round_loop(self, host: bool):
        """host=True: every peel round is driven by the host (PR_LOOP_HOST); False: the default, whole rounds queued on the
        device in score-all mode.  Same results either way."""
        _lib.check(self._L.plane_ransac_set_round_loop(self._h, 1 if host else 0))

    def flush_l2(self):
        _lib.check(self._L.plane_ransac_flush_l2(self._h))


# ---- host-side logic exported by the library (no device needed) ---------------------------------
def host_draw_triples(n_points: int, n_draws: int, seed: int = 12345) -> np.ndarray:
    out = np.empty((n_draws, 3), np.int32)
    _lib.check(_lib.load().plane_ransac_host_draw_triples(n_points, seed, n_draws, out.ctypes.data_as(C.c_void_p)))
    return out


def host_draw_triples_parallel(n_points: int, n_draws: int, seed: int = 12345):
    """The same triples through the parallel formulation of the sampler that the device-side round loop runs
    (csrc/pr_draw.h).  Returns None when it hands the round back to the sequential sampler (too many colliding ops)."""
    out = np.empty((n_draws, 3), np.int32)
    fb = C.c_int(0)
    _lib.check(_lib.load().plane_ransac_host_draw_triples_parallel(n_points, seed, n_draws, out.ctypes.data_as(C.c_void_p), C.byref(fb)))
    return None if fb.value else out


def host_replay(counts, good, n_points: int, max_iterations: int, probability: float):
    c = np.ascontiguousarray(counts, np.int32)
    g = np.ascontiguousarray(good, np.uint8)
    v = [C.c_int(0) for _ in range(5)]
    _lib.check(_lib.load().plane_ransac_host_replay(c.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), c.size,
                                                    n_points, max_iterations, probability, *[C.byref(x) for x in v]))
    return dict(best_draw=v[0].value, iterations=v[1].value, draws_used=v[2].value, skipped=v[3].value,
                exhausted=bool(v[4].value))


def host_rand_edges(seed: int, border_size: int) -> np.ndarray:
    """The ten border edges isPointInPoly draws after srand(seed) (MSVC CRT rand)."""
    out = np.zeros(10, np.int32)
    _lib.check(_lib.load().plane_ransac_host_rand_edges(int(seed) & 0xFFFFFFFF, int(border_size), out.ctypes.data_as(C.c_void_p)))
    return out


def host_shard_range(n_points: int, n_ranks: int, rank: int):
    a, b = C.c_longlong(0), C.c_longlong(0)
    _lib.check(_lib.load().plane_ransac_host_shard_range(n_points, n_ranks, rank, C.byref(a), C.byref(b)))
    return a.value, b.value


def host_plane_from_moments(moments, pivot, scale_exp: int) -> np.ndarray:
    m = np.ascontiguousarray(moments, np.int64)
    p = np.ascontiguousarray(pivot, np.float32)
    out = np.zeros(4, np.float32)
    _lib.check(_lib.load().plane_ransac_host_plane_from_moments(m.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
                                                                scale_exp, out.ctypes.data_as(C.c_void_p)))
    return out


def detect_planes(source_cloud: np.ndarray, distance_threshold: float = 0.1, max_iterations: int = 50,
                  min_plane_size: int = 500, device: int = 0, **kw):
    """PlaneDetect-style entry: cloud + (threshold, max iterations, minimum plane size) ->
    (list of Plane, remaining cloud).  The remaining cloud is what the reference leaves in
    source_cloud after postProcessPlanes (Dialog/PlaneDetect.h:1560-1566)."""
    with PlaneRansac(device) as pr:
        pr.set_cloud(source_cloud)
        ex = pr.extract_planes(make_params(distance_threshold, max_iterations, min_plane_size, **kw))
        return ex.planes, pr.remaining().copy()
