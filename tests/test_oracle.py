"""CPU tests of the oracle itself: the pins it does have (mt19937, analytic known answers, the
reference's double_shadow.pcd fixture) and its internal consistency.  The reference holds no tests or
golden outputs for this path, so PCL parity is unpinned; see oracle/pr_oracle.h."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_mt19937_known_answer(O):
    # C++11 [rand.predef]: the 10000th invocation of a default-constructed mt19937 is 4123659995
    assert int(O.mt19937_stream(5489, 10000)[-1]) == 4123659995


def test_mt19937_matches_numpy_legacy_seeding(O):
    bg = np.random.MT19937()
    bg._legacy_seeding(12345)  # init_genrand(12345) == boost::mt19937(12345u)
    assert (bg.random_raw(5000).astype(np.uint32) == O.mt19937_stream(12345, 5000)).all()


def test_draw_sequence_is_partial_fisher_yates(O):
    # replay drawIndexSample in Python from the raw stream: swap(shuffled[i], shuffled[i + (rng()>>1) % (N-i)])
    n, draws = 1000, 50
    raw = O.mt19937_stream(12345, 3 * draws).astype(np.int64) >> 1
    sh = list(range(n))
    exp = []
    for k in range(draws):
        for i in range(3):
            j = i + int(raw[3 * k + i]) % (n - i)
            sh[i], sh[j] = sh[j], sh[i]
        exp.append(sh[:3])
    got = O.draw_sequence(n, draws)
    assert (got == np.array(exp)).all()
    assert all(len(set(r)) == 3 for r in got.tolist())


def test_compute_model_known_answers(O):
    pts = np.array([[0, 0, 1], [1, 0, 1], [0, 1, 1], [2, 2, 2]], np.float32)
    ok, c = O.compute_model(pts, [0, 1, 2])
    assert ok and c.tolist() == [0.0, 0.0, 1.0, -1.0]
    ok, c = O.compute_model(pts, [0, 2, 1])  # opposite winding flips the normal
    assert ok and c.tolist() == [0.0, 0.0, -1.0, 1.0]
    # 3-4-5 triangle in an oblique plane: normal (3,0,4)/5 exactly representable after normalisation
    p = np.array([[0, 0, 0], [0, 1, 0], [4, 0, -3]], np.float32)
    ok, c = O.compute_model(p, [0, 1, 2])
    assert ok and np.allclose(c, [-0.6, 0.0, -0.8, 0.0], atol=1e-7)


def test_collinear_and_duplicate_samples_are_rejected(O):
    line = np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2], [1, 5, 2]], np.float32)
    assert not O.is_sample_good(line, [0, 1, 2])           # all three ratios equal
    assert not O.compute_model(line, [0, 1, 2])[0]
    assert O.is_sample_good(line, [0, 1, 3])
    dup = np.array([[1, 2, 3], [1, 2, 3], [4, 5, 7]], np.float32)
    assert not O.is_sample_good(dup, [0, 1, 2])            # p1 == p0: ratios 0,0,0
    assert not O.is_sample_good(dup, [0, 2, 1])           # p2 == p0: ratios +inf,+inf,+inf compare equal
    # three identical points give 0/0 = NaN in every component; NaN != NaN, so PCL calls the sample good
    same = np.ones((3, 3), np.float32)
    assert O.is_sample_good(same, [0, 1, 2])
    ok, c = O.compute_model(same, [0, 1, 2])
    assert ok and np.isnan(c).all()


def test_count_on_exact_plane(O):
    rng = np.random.default_rng(1)
    xy = rng.integers(-100, 100, size=(5000, 2)).astype(np.float32)
    on = np.c_[xy, np.full(5000, 2.0, np.float32)]
    off = np.c_[xy, np.full(5000, 2.5, np.float32)]
    pts = np.r_[on, off]
    coeff = np.array([0, 0, 1, -2], np.float32)
    for order in (O.DOT_PCL_SSE2, O.DOT_FMA):
        assert O.count_within(pts, coeff, 0.1, order) == 5000
        assert O.count_within(pts, coeff, 0.5, order) == 5000      # strict '<': |r| = 0.5 is out
        assert O.count_within(pts, coeff, 0.5000001, order) == 10000
        assert (O.select_within(pts, coeff, 0.1, order) == np.arange(5000)).all()
        assert O.count_within(pts, coeff, 0.1, order, mt=True) == 5000


def test_nan_and_inf_points_are_never_inliers(O):
    pts = np.array([[0, 0, 0], [np.nan, 0, 0], [0, np.inf, 0], [0, 0, -np.inf], [1, 1, 0]], np.float32)
    coeff = np.array([0, 0, 1, 0], np.float32)
    for order in (O.DOT_PCL_SSE2, O.DOT_FMA):
        assert O.select_within(pts, coeff, 0.1, order).tolist() == [0, 4]


def test_dot_orders_differ_only_inside_the_rounding_band(O, scene3):
    """The enumerated exception set of north_star: points whose decision differs between the PCL/SSE2
    order and the FMA order have ||r| - t| <= 1e-6 * (|ax|+|by|+|cz|+|d|) (two 4-term FP32 dot products
    differ by at most 2*gamma_4 ~ 4.8e-7 of that sum)."""
    pts = scene3.points(0, 400_000)
    tri = O.draw_sequence(pts.shape[0], 64)
    coeffs, good = O.models_from_triples(pts, tri)
    t = 0.1
    n_diff = 0
    for c in coeffs[good][:32]:
        r0 = O.residuals(pts, c, O.DOT_PCL_SSE2).astype(np.float64)
        r1 = O.residuals(pts, c, O.DOT_FMA).astype(np.float64)
        d = (np.abs(r0) < t) != (np.abs(r1) < t)
        mag = (np.abs(pts[:, :3].astype(np.float64) * c[:3].astype(np.float64)).sum(1) + abs(float(c[3])))
        assert np.all(np.abs(r0 - r1) <= 1e-6 * mag)
        band = np.abs(np.abs(r0) - t) <= 1e-6 * mag
        assert np.all(band[d]), "a decision flipped outside the stated band"
        n_diff += int(d.sum())
    # the two orders are genuinely different arithmetic (some residuals differ in the last ulp)
    assert n_diff >= 0


def _f64_plane(p):
    """Least-squares plane in float64 (numpy eigh) as an independent reference."""
    p = p.astype(np.float64)
    c = p.mean(0)
    w, v = np.linalg.eigh(np.cov((p - c).T, bias=True))
    n = v[:, 0]
    return np.r_[n, -n @ c]


def _align(a, ref):
    return -a if np.dot(a[:3], ref[:3]) < 0 else a


def test_refit_fixed_matches_float64_and_pcl_float_within_1e5(O, scene2):
    """north_star: plane coefficients within 1e-5 relative.  The order-independent integer-moment refit
    agrees with a float64 least-squares fit to ~1e-7; PCL's float accumulators agree with both to 1e-5
    on planes of a few thousand points and drift beyond it on large ones — that
    drift is PCL's own rounding, so 1e-5 is asserted against the float64 fit."""
    pts = scene2.points(0, 100_000)
    truth = scene2.patches[0].coeff.astype(np.float32)
    s = O.fixed_scale_exp(pts)
    for order in (O.DOT_PCL_SSE2, O.DOT_FMA):
        idx = O.select_within(pts, truth, 0.1, order)
        ref = _f64_plane(pts[idx, :3])
        b, mom = O.refit_fixed(pts, idx, pts[idx[0], :3], s, truth)
        b = _align(b.astype(np.float64), ref)
        assert np.abs(b - ref).max() <= 2e-7 * max(1.0, np.abs(ref).max())
        a_big = _align(O.refit_pcl_float(pts, idx, truth).astype(np.float64), ref)
        small = idx[:3000]
        ref_s = _f64_plane(pts[small, :3])
        a = _align(O.refit_pcl_float(pts, small, truth).astype(np.float64), ref_s)
        bs, _ = O.refit_fixed(pts, small, pts[small[0], :3], s, truth)
        bs = _align(bs.astype(np.float64), ref_s)
        assert np.abs(bs - ref_s).max() <= 2e-7 * max(1.0, np.abs(ref_s).max())
        assert np.abs(a - bs).max() <= 1e-5 * max(1.0, np.abs(ref_s).max())
        # on the full 30k-point plane PCL's float sums are the less accurate of the two
        assert np.abs(a_big - ref).max() >= np.abs(b - ref).max()
        # exact-arithmetic cross-check of the integer moments with Python ints
        sc = 2.0 ** s
        q = np.rint((pts[idx, :3].astype(np.float64) - pts[idx[0], :3].astype(np.float64)) * sc).astype(np.int64)
        tot = O.moments_total(mom)
        assert tot[0] == idx.size
        assert tot[1:4] == [int(q[:, a_].sum()) for a_ in range(3)]
        qq = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
        exact = [sum(int(u) * int(v) for u, v in zip(q[:, i].tolist(), q[:, j].tolist())) for i, j in qq]
        assert tot[4:] == exact
        assert np.abs(q).max() < 2 ** 30


def test_refit_fixed_is_order_independent(O, scene2):
    pts = scene2.points(0, 50_000)
    truth = scene2.patches[1].coeff.astype(np.float32)
    idx = O.select_within(pts, truth, 0.1, O.DOT_FMA)
    s = O.fixed_scale_exp(pts)
    piv = pts[idx[0], :3]
    a, ma = O.refit_fixed(pts, idx, piv, s, truth)
    rng = np.random.default_rng(0)
    b, mb = O.refit_fixed(pts, rng.permutation(idx), piv, s, truth)
    assert a.tobytes() == b.tobytes() and (ma == mb).all()


def test_refit_with_fewer_than_four_inliers_returns_input(O):
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    cin = np.array([0, 0, 1, 0], np.float32)
    assert O.refit_pcl_float(pts, [0, 1, 2], cin).tolist() == cin.tolist()
    out, _ = O.refit_fixed(pts, [0, 1, 2], pts[0], 0, cin)
    assert out.tolist() == cin.tolist()


def test_segment_finds_the_dominant_plane(O, scene2):
    pts = scene2.points(0, 60_000)
    prm = O.make_params(0.1, 255, 500, 1.0, True, 12345, 8, O.DOT_FMA, O.REFIT_FIXED)
    seg = O.segment(pts, prm)
    assert seg.ok and seg.trace.iterations == 256 and seg.trace.draws >= 256
    truth = scene2.patches[0].coeff
    c = seg.coeff.astype(np.float64)
    if np.dot(c[:3], truth[:3]) < 0:
        c = -c
    assert np.abs(c - truth).max() < 0.05
    assert (np.diff(seg.inliers) > 0).all()
    assert seg.inliers.size == O.count_within(pts, seg.coeff, 0.1, O.DOT_FMA)


def test_adaptive_exit_and_iteration_cap(O, scene2):
    pts = scene2.points(0, 30_000)
    seg = O.segment(pts, O.make_params(0.1, 50, 500, 0.99, True))
    assert seg.ok and seg.trace.iterations == 51          # w ~ 0.35: k ~ 105 > cap -> max_iterations + 1 trials
    flat = np.c_[np.random.default_rng(0).random((5000, 2)), np.zeros(5000)].astype(np.float32)
    seg = O.segment(flat, O.make_params(0.1, 50, 500, 0.99, True))
    assert seg.ok and seg.trace.iterations == 1 and seg.inliers.size == 5000   # w = 1: first trial ends it


def test_segment_degenerate_inputs(O):
    prm = O.make_params(0.1, 50, 1, 0.99, True)
    assert not O.segment(np.zeros((0, 3), np.float32), prm).ok
    assert not O.segment(np.zeros((2, 3), np.float32), prm).ok
    i = np.arange(1, 51, dtype=np.float32)
    line = np.c_[i, 2 * i, 4 * i]                 # every sample is collinear: 1000 redraws, then "no samples"
    seg = O.segment(line, prm)
    assert not seg.ok and seg.trace.draws == 1000 and seg.inliers.size == 0
    same = np.ones((50, 3), np.float32)           # 0/0 ratios pass isSampleGood; the model is NaN, 0 inliers
    seg = O.segment(same, prm)
    assert seg.ok and np.isnan(seg.coeff).all() and seg.inliers.size == 0 and seg.trace.iterations == 51


def test_extract_planes_peels_in_order(O, scene2):
    pts = scene2.points(0, 40_000)
    prm = O.make_params(0.1, 255, 2000, 1.0, True, 12345, 8, O.DOT_FMA, O.REFIT_FIXED)
    ex = O.extract_planes(pts, prm)
    assert len(ex.coeffs) == 3
    sizes = [len(i) for i in ex.inliers_cur]
    assert sizes == sorted(sizes, reverse=True)
    # partition: every original index is in exactly one plane or in the remaining cloud, order preserved
    allidx = np.concatenate(ex.inliers_orig)
    assert len(set(allidx.tolist())) == allidx.size
    rest = np.setdiff1d(np.arange(pts.shape[0]), allidx)
    assert ex.remaining.shape[0] == rest.size and (ex.remaining == pts[rest]).all()
    # inliers_cur of round 0 are original indices
    assert (ex.inliers_cur[0] == ex.inliers_orig[0]).all()


@pytest.mark.parametrize("name", ["double_shadow_golden.json", "synthetic_golden.json"])
def test_golden_fixtures(O, double_shadow, scene2, name):
    cases = json.load(open(os.path.join(GOLDEN, name)))
    pts = double_shadow if name.startswith("double") else scene2.points(0, 20000)
    for c in cases:
        seg = O.segment(pts, O.make_params(**c["params"]))
        assert seg.ok == c["ok"]
        assert [float(v).hex() for v in seg.coeff] == c["coeff"]
        assert seg.inliers.size == c["n_inliers"]
        assert seg.trace.iterations == c["iterations"] and seg.trace.draws == c["draws"]
        assert list(seg.trace.best_sample) == c["best_sample"] and seg.trace.best_count == c["best_count"]


def test_double_shadow_reference_threshold_is_degenerate(O, double_shadow):
    # Dialog/config.txt:29 T_dist_point_plane = 0.1 swallows the whole 0.15 m blob: first hypothesis wins
    seg = O.segment(double_shadow, O.make_params(0.1, 50, 500, 0.99, True, dot_order=O.DOT_PCL_SSE2,
                                                 refit_mode=O.REFIT_PCL_FLOAT))
    assert seg.ok and seg.inliers.size == 991 and seg.trace.iterations == 1


def test_line_model_replays_the_reference_sac_call_on_its_saved_contours(O):
    """The reference's only literal pcl::SACSegmentation call (Dialog/SimplifyVerticesSize.cpp:64-67,87,122,146): SACMODEL_LINE,
    SAC_RANSAC, setDistanceThreshold(FLT_MAX), defaults otherwise, on a growing run of contour vertices.  Replayed here on
    the 11 plane borders the reference saved (Dialog/dataForPlane/source_plane_registration.pcd, tests/golden/
    ref_polygons_golden.npz) through the SAME sampler and computeModel loop the plane path uses — a second, independent
    consumer of that code with the reference's own call pattern: every vertex is an inlier of the first good 2-point
    sample, w = 1 makes k collapse, and the loop ends after one iteration."""
    g = np.load(os.path.join(GOLDEN, "ref_polygons_golden.npz"))
    verts, sizes = g["vertices"], g["sizes"]
    flt_max = float(np.finfo(np.float32).max)
    prm = O.make_params(flt_max, 50, 0, 0.99, True, 12345, 1, O.DOT_FMA, O.REFIT_PCL_FLOAT)
    start, fits = 0, 0
    for poly, m in enumerate(sizes.tolist()):
        contour = verts[start:start + m]
        start += m
        for seed in range(0, m, max(1, m // 6)):
            for grow in (1, 2, 4, 9):                      # prev / seed / next, then the run as constructInitLineSegs grows it
                idx = [(seed + d) % m for d in range(-grow, grow + 1)]
                run = np.ones((len(idx), 4), np.float32)
                run[:, :3] = contour[idx]
                n = run.shape[0]
                ok, coeff, inl, tr = O.segment_line(run, prm)
                # the sampler, restated independently: partial Fisher-Yates with a 2-point sample over mt19937(12345) >> 1
                raw = O.mt19937_stream(12345, 2 * 1000).astype(np.int64) >> 1
                sh, draws, pair = list(range(n)), 0, None
                for k in range(1000):
                    for i in range(2):
                        j = i + int(raw[2 * k + i]) % (n - i)
                        sh[i], sh[j] = sh[j], sh[i]
                    draws += 1
                    a, b = run[sh[0], :3], run[sh[1], :3]
                    if (a != b).all():                     # PCL 1.8 isSampleGood: x, y AND z differ
                        pair = sh[:2]
                        break
                assert ok == (pair is not None)
                if not ok:
                    continue
                assert tr.iterations == 1 and tr.draws == draws and list(tr.best_sample)[:2] == pair
                assert tr.best_count == n and inl.size == n and (inl == np.arange(n)).all()
                # optimizeModelCoefficients: point = centroid, direction = principal axis (float64 PCA as the yardstick)
                p64 = run[:, :3].astype(np.float64)
                assert np.abs(coeff[:3] - p64.mean(0)).max() <= 1e-6 * max(1.0, np.abs(p64).max())
                w, v = np.linalg.eigh(np.cov((p64 - p64.mean(0)).T))
                if w[2] > 1e-9 and w[2] > 4 * w[1]:        # a run that is a line, not a corner: the axis is well defined
                    assert abs(float(coeff[3:] @ v[:, 2])) >= 1 - 1e-5
                assert abs(float(np.linalg.norm(coeff[3:])) - 1) <= 1e-6
                fits += 1
    assert fits > 200, fits
    print(f'line-model fits replayed: {fits}')
    # the k-point sampler against the same Python walk on a larger index set
    raw = O.mt19937_stream(12345, 2 * 40).astype(np.int64) >> 1
    sh, exp = list(range(997)), []
    for k in range(40):
        for i in range(2):
            j = i + int(raw[2 * k + i]) % (997 - i)
            sh[i], sh[j] = sh[j], sh[i]
        exp.append(sh[:2])
    assert (O.draw_sequence_k(997, 40, 2) == np.array(exp)).all()
    assert (O.draw_sequence_k(500, 30, 3) == O.draw_sequence(500, 30)).all()
