"""ctypes view of the CPU oracle (oracle/pr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; nothing under dialog_b200/ does.  The oracle restates PCL 1.8's SACSegmentation(SACMODEL_PLANE)
+ ExtractIndices behaviour (parity unpinned by the reference — see pr_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpr_oracle.so")

DOT_PCL_SSE2 = 0
DOT_FMA = 1
REFIT_PCL_FLOAT = 0
REFIT_FIXED = 1
INT_MIN = -(2**31)


class Params(C.Structure):
    _fields_ = [
        ("distance_threshold", C.c_double),
        ("max_iterations", C.c_int),
        ("min_plane_size", C.c_int),
        ("probability", C.c_double),
        ("optimize_coefficients", C.c_int),
        ("seed", C.c_uint),
        ("max_planes", C.c_int),
        ("dot_order", C.c_int),
        ("refit_mode", C.c_int),
    ]


class Trace(C.Structure):
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
    ]


def make_params(distance_threshold=0.1, max_iterations=50, min_plane_size=500, probability=0.99,
                optimize_coefficients=True, seed=12345, max_planes=64, dot_order=DOT_FMA,
                refit_mode=REFIT_FIXED) -> Params:
    return Params(float(distance_threshold), int(max_iterations), int(min_plane_size), float(probability),
                  int(bool(optimize_coefficients)), int(seed), int(max_planes), int(dot_order),
                  int(refit_mode))


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc -O2 -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("pr_oracle.c", "pr_oracle_poly.c", "pr_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
        L.orc_draw_sequence.argtypes = [C.c_size_t, C.c_uint32, C.c_int, vp]
        L.orc_is_sample_good.argtypes = [vp, vp]
        L.orc_compute_model.argtypes = [vp, vp, vp]
        L.orc_residuals.argtypes = [vp, C.c_size_t, vp, C.c_int, vp]
        L.orc_count_within.argtypes = [vp, C.c_size_t, vp, C.c_double, C.c_int]
        L.orc_count_within.restype = C.c_int64
        L.orc_count_within_mt.argtypes = [vp, C.c_size_t, vp, C.c_double, C.c_int]
        L.orc_count_within_mt.restype = C.c_int64
        L.orc_select_within.argtypes = [vp, C.c_size_t, vp, C.c_double, C.c_int, vp]
        L.orc_select_within.restype = C.c_size_t
        L.orc_count_batch.argtypes = [vp, C.c_size_t, vp, C.c_int, C.c_double, C.c_int, C.c_int, vp]
        L.orc_count_batch.restype = None
        L.orc_refit_pcl_float.argtypes = [vp, vp, C.c_size_t, vp, vp]
        L.orc_fixed_scale_exp.argtypes = [vp, C.c_size_t]
        L.orc_refit_fixed.argtypes = [vp, vp, C.c_size_t, vp, C.c_int, vp, vp, vp]
        L.orc_plane_from_moments.argtypes = [vp, vp, C.c_int, vp]
        L.orc_segment.argtypes = [vp, C.c_size_t, C.POINTER(Params), C.c_int, vp, vp,
                                  C.POINTER(C.c_size_t), C.POINTER(Trace)]
        L.orc_extract_planes.argtypes = [vp, C.c_size_t, C.POINTER(Params), vp, vp, vp, C.c_size_t, vp,
                                         C.POINTER(C.c_int), vp, C.POINTER(C.c_size_t), C.POINTER(Trace)]
        L.orc_remove_nonfinite.argtypes = [vp, C.c_size_t, vp, vp]
        L.orc_remove_nonfinite.restype = C.c_size_t
        L.orc_centroid_exact.argtypes = [vp, C.c_size_t, vp]
        L.orc_translate.argtypes = [vp, C.c_size_t, vp]
        L.orc_translate.restype = None
        L.orc_project_points.argtypes = [vp, vp, C.c_size_t, vp, vp]
        L.orc_project_points.restype = None
        L.orc_msvc_rand_edges.argtypes = [C.c_uint, C.c_int, vp]
        L.orc_msvc_rand_edges.restype = None
        L.orc_segs_intersect.argtypes = [vp, vp, vp, vp]
        L.orc_point_in_poly.argtypes = [vp, vp, vp, C.c_int, C.c_float, C.c_uint]
        L.orc_reabsorb.argtypes = [vp, C.c_size_t, vp, vp, vp, C.c_int, C.c_float, C.c_uint, vp, C.c_size_t, vp, vp,
                                   C.POINTER(C.c_size_t)]
        L.orc_estimate_normals.argtypes = [vp, C.c_size_t, C.c_double, vp, C.c_int, vp, vp]
        L.orc_normals_scale_exp.argtypes = [C.c_double]
        L.orc_cluster_filter.argtypes = [vp, C.c_size_t, C.c_double, C.c_int, vp]
        L.orc_mt_seed.argtypes = [vp, C.c_uint32]
        L.orc_mt_next.argtypes = [vp]
        L.orc_mt_next.restype = C.c_uint32
        _lib = L
    return _lib


def _cloud(a: np.ndarray) -> np.ndarray:
    """(N,4) float32 C-contiguous view of an (N,3|4) array (w = 1, like pcl::PointXYZ)."""
    a = np.asarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError("cloud must be (N,3) or (N,4)")
    if a.shape[1] == 3:
        b = np.ones((a.shape[0], 4), np.float32)
        b[:, :3] = a
        a = b
    return np.ascontiguousarray(a)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def mt19937_stream(seed: int, n: int) -> np.ndarray:
    st = (C.c_uint32 * 625)()
    L = lib()
    L.orc_mt_seed(st, seed)
    return np.array([L.orc_mt_next(st) for _ in range(n)], dtype=np.uint32)


def draw_sequence(n_points: int, n_draws: int, seed: int = 12345) -> np.ndarray:
    out = np.empty((n_draws, 3), np.int32)
    if lib().orc_draw_sequence(n_points, seed, n_draws, _p(out)) != 0:
        raise ValueError("cloud too small to sample")
    return out


def is_sample_good(cloud, idx) -> bool:
    c = _cloud(cloud)
    i = np.asarray(idx, np.int32)
    return bool(lib().orc_is_sample_good(_p(c), _p(i)))


def compute_model(cloud, idx):
    c = _cloud(cloud)
    i = np.ascontiguousarray(idx, np.int32)
    out = np.zeros(4, np.float32)
    ok = lib().orc_compute_model(_p(c), _p(i), _p(out))
    return bool(ok), out


def models_from_triples(cloud, triples):
    """coeffs (K,4) float32 (NaN rows for degenerate samples) and good flags (K,)."""
    c = _cloud(cloud)
    t = np.ascontiguousarray(triples, np.int32).reshape(-1, 3)
    coeffs = np.full((t.shape[0], 4), np.nan, np.float32)
    good = np.zeros(t.shape[0], bool)
    L = lib()
    tmp = np.zeros(4, np.float32)
    for k in range(t.shape[0]):
        row = np.ascontiguousarray(t[k])
        if L.orc_is_sample_good(_p(c), _p(row)) and L.orc_compute_model(_p(c), _p(row), _p(tmp)):
            coeffs[k] = tmp
            good[k] = True
    return coeffs, good


def residuals(cloud, coeff, dot_order=DOT_FMA) -> np.ndarray:
    c = _cloud(cloud)
    co = np.ascontiguousarray(coeff, np.float32)
    out = np.empty(c.shape[0], np.float32)
    lib().orc_residuals(_p(c), c.shape[0], _p(co), dot_order, _p(out))
    return out


def count_within(cloud, coeff, t, dot_order=DOT_FMA, mt=False) -> int:
    c = _cloud(cloud)
    co = np.ascontiguousarray(coeff, np.float32)
    f = lib().orc_count_within_mt if mt else lib().orc_count_within
    return int(f(_p(c), c.shape[0], _p(co), float(t), dot_order))


def count_batch(cloud, coeffs, t, dot_order=DOT_FMA, threads=1) -> np.ndarray:
    c = _cloud(cloud)
    co = np.ascontiguousarray(coeffs, np.float32).reshape(-1, 4)
    out = np.zeros(co.shape[0], np.int32)
    lib().orc_count_batch(_p(c), c.shape[0], _p(co), co.shape[0], float(t), dot_order, threads, _p(out))
    return out


def select_within(cloud, coeff, t, dot_order=DOT_FMA) -> np.ndarray:
    c = _cloud(cloud)
    co = np.ascontiguousarray(coeff, np.float32)
    out = np.empty(c.shape[0], np.int32)
    m = lib().orc_select_within(_p(c), c.shape[0], _p(co), float(t), dot_order, _p(out))
    return out[:m].copy()


def refit_pcl_float(cloud, idx, coeff_in) -> np.ndarray:
    c = _cloud(cloud)
    i = np.ascontiguousarray(idx, np.int32)
    ci = np.ascontiguousarray(coeff_in, np.float32)
    out = np.zeros(4, np.float32)
    lib().orc_refit_pcl_float(_p(c), _p(i), i.size, _p(ci), _p(out))
    return out


def fixed_scale_exp(cloud) -> int:
    c = _cloud(cloud)
    return int(lib().orc_fixed_scale_exp(_p(c), c.shape[0]))


def refit_fixed(cloud, idx, pivot, scale_exp, coeff_in):
    c = _cloud(cloud)
    i = np.ascontiguousarray(idx, np.int32)
    pv = np.ascontiguousarray(pivot, np.float32)
    ci = np.ascontiguousarray(coeff_in, np.float32)
    out = np.zeros(4, np.float32)
    mom = np.zeros(16, np.int64)
    lib().orc_refit_fixed(_p(c), _p(i), i.size, _p(pv), scale_exp, _p(ci), _p(out), _p(mom))
    return out, mom


def plane_from_moments(moments, pivot, scale_exp):
    m = np.ascontiguousarray(moments, np.int64)
    pv = np.ascontiguousarray(pivot, np.float32)
    out = np.zeros(4, np.float32)
    ok = lib().orc_plane_from_moments(_p(m), _p(pv), scale_exp, _p(out))
    return bool(ok), out


def moments_total(m: np.ndarray):
    """Recombine the (hi, lo) split second moments into Python ints: [n, Sx, Sy, Sz, Sxx..Szz]."""
    m = [int(v) for v in m]
    return m[:4] + [m[4 + 2 * k] * (1 << 32) + m[5 + 2 * k] for k in range(6)]


def preprocess(cloud, remove_nonfinite=True, translate=True):
    """preProcess's first two steps: (filtered + translated cloud, source index map, centroid)."""
    c = _cloud(cloud).copy()
    idx = np.arange(c.shape[0], dtype=np.int32)
    if remove_nonfinite:
        out = np.empty_like(c)
        m = lib().orc_remove_nonfinite(_p(c), c.shape[0], _p(out), _p(idx))
        c, idx = out[:m].copy(), idx[:m].copy()
    cen = np.zeros(3, np.float32)
    if translate and lib().orc_centroid_exact(_p(c), c.shape[0], _p(cen)):
        lib().orc_translate(_p(c), c.shape[0], _p(cen))
    return c, idx, cen


def project_points(cloud, idx, coeff) -> np.ndarray:
    c = _cloud(cloud)
    i = np.ascontiguousarray(idx, np.int32)
    co = np.ascontiguousarray(coeff, np.float32)
    out = np.empty((i.size, 4), np.float32)
    lib().orc_project_points(_p(c), _p(i), i.size, _p(co), _p(out))
    return out


@dataclass
class Segment:
    ok: bool
    coeff: np.ndarray
    inliers: np.ndarray
    trace: Trace


def segment(cloud, params: Params, scale_exp: int = INT_MIN) -> Segment:
    c = _cloud(cloud)
    coeff = np.zeros(4, np.float32)
    inl = np.empty(max(c.shape[0], 1), np.int32)
    n = C.c_size_t(0)
    tr = Trace()
    ok = lib().orc_segment(_p(c), c.shape[0], C.byref(params), scale_exp, _p(coeff), _p(inl), C.byref(n),
                           C.byref(tr))
    return Segment(ok == 1, coeff, inl[: n.value].copy(), tr)


@dataclass
class Extraction:
    coeffs: np.ndarray          # (P,4)
    inliers_cur: list           # per plane: indices into that round's cloud
    inliers_orig: list          # per plane: indices into the input cloud
    remaining: np.ndarray       # (R,4)
    traces: list


def extract_planes(cloud, params: Params) -> Extraction:
    c = _cloud(cloud)
    n = c.shape[0]
    mp = params.max_planes
    coeffs = np.zeros((mp, 4), np.float32)
    cur = np.empty(max(n, 1), np.int32)
    orig = np.empty(max(n, 1), np.int32)
    offs = np.zeros(mp + 1, np.uintp)
    npl = C.c_int(0)
    rem = np.empty((max(n, 1), 4), np.float32)
    nrem = C.c_size_t(0)
    traces = (Trace * (mp + 1))()
    rc = lib().orc_extract_planes(_p(c), n, C.byref(params), _p(coeffs), _p(cur), _p(orig), n, _p(offs),
                                  C.byref(npl), _p(rem), C.byref(nrem), traces)
    if rc != 0:
        raise RuntimeError(f"orc_extract_planes failed: {rc}")
    P = npl.value
    o = [int(v) for v in offs[: P + 1]]
    return Extraction(coeffs[:P].copy(), [cur[o[k]: o[k + 1]].copy() for k in range(P)],
                      [orig[o[k]: o[k + 1]].copy() for k in range(P)], rem[: nrem.value].copy(),
                      [traces[k] for k in range(min(P + 1, mp))])


# ---- postProcessPlanes re-absorption (Dialog/PlaneDetect.h:1454-1580) ---------------------------------------
def msvc_rand_edges(seed: int, border_size: int) -> np.ndarray:
    out = np.zeros(10, np.int32)
    lib().orc_msvc_rand_edges(int(seed) & 0xFFFFFFFF, int(border_size), _p(out))
    return out


def segs_intersect(pa, pb, pc, pd) -> bool:
    a, b, c, d = (_cloud(np.asarray(v, np.float32).reshape(1, -1)) for v in (pa, pb, pc, pd))
    return bool(lib().orc_segs_intersect(_p(a), _p(b), _p(c), _p(d)))


def points_in_poly(points, coeff, border, t, seed) -> np.ndarray:
    pts, bd = _cloud(points), _cloud(border)
    co = np.ascontiguousarray(coeff, np.float32)
    L = lib()
    out = np.zeros(pts.shape[0], bool)
    for i in range(pts.shape[0]):
        out[i] = bool(L.orc_point_in_poly(C.c_void_p(pts.ctypes.data + 16 * i), _p(co), _p(bd), bd.shape[0], float(t),
                                          int(seed) & 0xFFFFFFFF))
    return out


@dataclass
class Reabsorption:
    absorbed: list              # per plane: ascending indices of the points it claimed
    remaining_idx: np.ndarray   # ascending indices of the unclaimed points


def reabsorb(cloud, coeffs, borders, t, seed) -> Reabsorption:
    """borders: list of (nb_j, 3|4) arrays, one polygon per plane."""
    c = _cloud(cloud)
    co = np.ascontiguousarray(coeffs, np.float32).reshape(-1, 4)
    P = co.shape[0]
    bd = np.concatenate([_cloud(b) for b in borders]) if P else np.zeros((0, 4), np.float32)
    offs = np.zeros(P + 1, np.uintp)
    offs[1:] = np.cumsum([len(b) for b in borders])
    cap = max(1, c.shape[0] * max(P, 1))
    ab = np.empty(cap, np.int32)
    po = np.zeros(P + 1, np.uintp)
    rem = np.empty(max(1, c.shape[0]), np.int32)
    nrem = C.c_size_t(0)
    rc = lib().orc_reabsorb(_p(c), c.shape[0], _p(co), _p(bd), _p(offs), P, float(t), int(seed) & 0xFFFFFFFF, _p(ab), cap,
                            _p(po), _p(rem), C.byref(nrem))
    if rc != 0:
        raise RuntimeError("orc_reabsorb failed (empty border?)")
    o = [int(v) for v in po]
    return Reabsorption([ab[o[j]: o[j + 1]].copy() for j in range(P)], rem[: nrem.value].copy())


NORMALS_PCL_FLOAT = 0
NORMALS_FIXED = 1


def estimate_normals(cloud, radius, viewpoint=(0.0, 0.0, 0.0), mode=NORMALS_FIXED):
    """pcl::NormalEstimationOMP with a radius search: ((n,4) normal_x, normal_y, normal_z, curvature; neighbour counts)."""
    c = _cloud(cloud)
    vp = np.ascontiguousarray(viewpoint, np.float32)
    out = np.empty((c.shape[0], 4), np.float32)
    cnt = np.zeros(c.shape[0], np.int32)
    if lib().orc_estimate_normals(_p(c), c.shape[0], float(radius), _p(vp), int(mode), _p(out), _p(cnt)) != 0:
        raise ValueError("bad radius")
    return out, cnt


def cluster_filter(cloud, radius, max_small_cluster) -> np.ndarray:
    """clusterFilt: boolean keep mask (False for points of radius-graph components with <= max_small_cluster points)."""
    c = _cloud(cloud)
    keep = np.ones(c.shape[0], np.uint8)
    if lib().orc_cluster_filter(_p(c), c.shape[0], float(radius), int(max_small_cluster), _p(keep)) != 0:
        raise ValueError("bad radius")
    return keep.astype(bool)


def segment_line(cloud, params: Params):
    """pcl::SACSegmentation with SACMODEL_LINE as Dialog/SimplifyVerticesSize.cpp:64-67 calls it: (ok, coeff[6] = point +
    direction, inlier indices, trace) through the same sampler and computeModel loop as the plane model."""
    c = _cloud(cloud)
    coeff = np.zeros(6, np.float32)
    inl = np.empty(max(c.shape[0], 1), np.int32)
    n = C.c_size_t(0)
    tr = Trace()
    L = lib()
    L.orc_segment_line.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(Params), C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(Trace)]
    rc = L.orc_segment_line(_p(c), c.shape[0], C.byref(params), _p(coeff), _p(inl), C.byref(n), C.byref(tr))
    if rc < 0:
        raise RuntimeError("orc_segment_line failed")
    return bool(rc), coeff, inl[: n.value].copy(), tr


def draw_sequence_k(n_points: int, n_draws: int, sample_size: int, seed: int = 12345) -> np.ndarray:
    """drawIndexSample for a model of sample_size points (line: 2), through the oracle's sampler."""
    L = lib()
    class Sampler(C.Structure):
        _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int), ("shuffled", C.c_void_p), ("n", C.c_size_t)]
    s = Sampler()
    L.orc_sampler_init.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
    L.orc_sampler_draw_k.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_sampler_free.argtypes = [C.c_void_p]
    if L.orc_sampler_init(C.byref(s), n_points, seed):
        raise MemoryError
    out = np.zeros((n_draws, sample_size), np.int32)
    for k in range(n_draws):
        L.orc_sampler_draw_k(C.byref(s), sample_size, out[k].ctypes.data_as(C.c_void_p))
    L.orc_sampler_free(C.byref(s))
    return out


def cluster_filter_reference_bfs(cloud, radius, max_small_cluster, ties_by_index: bool) -> np.ndarray:
    """clusterFilt restated literally (BFS, first radius-search hit skipped, Dialog/PlaneDetect.h:1598-1634): keep mask."""
    c = _cloud(cloud)
    keep = np.ones(c.shape[0], np.uint8)
    L = lib()
    L.orc_cluster_filter_reference_bfs.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p]
    if L.orc_cluster_filter_reference_bfs(_p(c), c.shape[0], float(radius), int(max_small_cluster), int(ties_by_index), _p(keep)) != 0:
        raise ValueError("bad radius")
    return keep.astype(bool)


# ---- the reference's own source of the same predicate (oracle/build_ref.py) ------------------------------------
_REF_SO = os.path.join(_HERE, "_ref", "libdialog_ref.so")
_ref = None


def ref_lib():
    """oracle/_ref/libdialog_ref.so or None.  Built from /root/reference where that exists; prebuilt otherwise."""
    global _ref
    if _ref is None:
        if os.path.exists("/root/reference/Dialog/PlaneDetect.h"):
            import importlib.util
            spec = importlib.util.spec_from_file_location("_orc_build_ref", os.path.join(_HERE, "build_ref.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        if not os.path.exists(_REF_SO):
            return None
        L = C.CDLL(_REF_SO)
        vp = C.c_void_p
        L.ref_segs_intersect.argtypes = [vp, vp, vp, vp]
        L.ref_points_in_poly.argtypes = [vp, C.c_size_t, vp, vp, C.c_int, C.c_float, C.c_uint, vp]
        L.ref_points_in_poly.restype = None
        L.ref_project.argtypes = [vp, vp, vp, vp]
        L.ref_project.restype = None
        _ref = L
    return _ref


def ref_points_in_poly(points, coeff, border, t, seed) -> np.ndarray:
    pts, bd = _cloud(points), _cloud(border)
    co = np.ascontiguousarray(coeff, np.float32)
    out = np.zeros(pts.shape[0], np.uint8)
    ref_lib().ref_points_in_poly(_p(pts), pts.shape[0], _p(co), _p(bd), bd.shape[0], float(t), int(seed) & 0xFFFFFFFF, _p(out))
    return out.astype(bool)


def ref_segs_intersect(pa, pb, pc, pd) -> bool:
    a, b, c, d = (np.ascontiguousarray(v, np.float32)[:3].copy() for v in (pa, pb, pc, pd))
    return bool(ref_lib().ref_segs_intersect(_p(a), _p(b), _p(c), _p(d)))


def ref_project(p, coeff):
    pp = np.ascontiguousarray(p, np.float32)[:3].copy()
    co = np.ascontiguousarray(coeff, np.float32)
    out = np.zeros(3, np.float32)
    d = np.zeros(1, np.float32)
    ref_lib().ref_project(_p(pp), _p(co), _p(out), _p(d))
    return out, float(d[0])
