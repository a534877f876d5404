"""postProcessPlanes re-absorption (Dialog/PlaneDetect.h:1454-1580).

CPU part: the oracle restatement (oracle/pr_oracle_poly.c) against the reference's OWN source of isPointInPoly /
isBothLineSegsIntersect / projPoint2Plane compiled over a type shim (oracle/build_ref.py -> oracle/_ref), and against
the committed golden vectors generated from it (tests/golden/make_reabsorb_golden.py).
GPU part: plane_ransac_reabsorb through the C ABI against the oracle, bit-exact index lists and remaining cloud.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reabsorb_golden.npz")


def make_plane(rng, nb, R=1.0, centre_scale=1.0):
    """A star-shaped polygon with nb vertices in a random oblique plane: (coeff, border (nb,4), frame)."""
    n = rng.normal(size=3)
    n /= np.linalg.norm(n)
    u = np.cross(n, [1.0, 0.0, 0.0])
    if np.linalg.norm(u) < 1e-3:
        u = np.cross(n, [0.0, 1.0, 0.0])
    u /= np.linalg.norm(u)
    v = np.cross(n, u)
    o = rng.uniform(-1, 1, 3) * centre_scale
    ang = np.sort(rng.uniform(0, 2 * np.pi, nb))
    r = R * (0.6 + 0.4 * rng.uniform(size=nb))
    bd = o + np.outer(r * np.cos(ang), u) + np.outer(r * np.sin(ang), v)
    n32 = n.astype(np.float32)
    coeff = np.array([n32[0], n32[1], n32[2], -np.float32(n @ o)], np.float32)
    b4 = np.ones((nb, 4), np.float32)
    b4[:, :3] = bd
    return coeff, b4, (o, u, v, n)


def points_near(rng, frame, m, spread=1.3, sigma=0.08):
    o, u, v, n = frame
    ab = rng.uniform(-spread, spread, (m, 2))
    h = rng.normal(0, sigma, m)
    p = np.ones((m, 4), np.float32)
    p[:, :3] = o + np.outer(ab[:, 0], u) + np.outer(ab[:, 1], v) + np.outer(h, n)
    return p


def scene(seed, n_planes=5, pts_per_plane=3000, n_clutter=5000):
    rng = np.random.default_rng(seed)
    coeffs, borders, parts = [], [], []
    for _ in range(n_planes):
        c, b, fr = make_plane(rng, int(rng.integers(3, 90)))
        coeffs.append(c)
        borders.append(b)
        parts.append(points_near(rng, fr, pts_per_plane))
    clutter = np.ones((n_clutter, 4), np.float32)
    clutter[:, :3] = rng.uniform(-2.5, 2.5, (n_clutter, 3))
    parts.append(clutter)
    cloud = np.concatenate(parts)
    cloud = cloud[rng.permutation(len(cloud))]
    return np.ascontiguousarray(cloud), np.array(coeffs, np.float32), borders


# ---------------------------------------------------------------------------------------------------------------
# CPU: oracle vs the reference's own source, vs golden vectors, and known answers
# ---------------------------------------------------------------------------------------------------------------
def test_msvc_rand_known_answers():
    # MSVC CRT: srand(1) -> 41, 18467, 6334, 26500, 19169, 15724, 11478, 29358, 26962, 24464
    first = [41, 18467, 6334, 26500, 19169, 15724, 11478, 29358, 26962, 24464]
    assert O.msvc_rand_edges(1, 32768).tolist() == first
    assert O.msvc_rand_edges(1, 7).tolist() == [v % 7 for v in first]


def test_square_known_answers():
    sq = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    coeff = np.array([0, 0, 1, 0], np.float32)
    pts = np.array([[0.5, 0.5, 0.0], [0.5, 0.5, 0.05], [0.5, 0.5, 0.1], [0.5, 0.5, 0.11], [1.5, 0.5, 0.0], [-0.2, 0.3, 0.01],
                    [0.25, 0.75, -0.09], [np.nan, 0.5, 0.0]], np.float32)
    for seed in (0, 1, 12345, 2**31 - 1):
        got = O.points_in_poly(pts, coeff, sq, 0.1, seed)
        # dist == T stays a candidate (the reference rejects only dist > T); 0.11 is out; NaN is never inside
        assert got.tolist() == [True, True, True, False, False, False, True, False]


def test_segs_intersect_known_answers():
    assert O.segs_intersect([0, 0, 0], [2, 0, 0], [1, -1, 0], [1, 1, 0])
    assert not O.segs_intersect([0, 0, 0], [2, 0, 0], [3, -1, 0], [3, 1, 0])      # crosses the line beyond b
    assert not O.segs_intersect([0, 0, 0], [2, 0, 0], [0, 1, 0], [2, 1, 0])       # parallel: n.n >= 0.9999
    assert not O.segs_intersect([0, 0, 0], [0, 0, 0], [1, -1, 0], [1, 1, 0])      # zero-length edge


@pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_reference_source():
    rng = np.random.default_rng(7)
    total = 0
    for trial in range(12):
        coeff, bd, fr = make_plane(rng, int(rng.integers(3, 120)), centre_scale=[1.0, 30.0][trial % 2])
        pts = points_near(rng, fr, 3000)
        seed = int(rng.integers(0, 2**32))
        a = O.points_in_poly(pts, coeff, bd, 0.1, seed)
        b = O.ref_points_in_poly(pts, coeff, bd, 0.1, seed)
        assert np.array_equal(a, b)
        total += int(a.sum())
        for i in range(0, 3000, 500):
            pr, dist = O.ref_project(pts[i], coeff)
            assert np.array_equal(O.project_points(pts, [i], coeff)[0, :3].view(np.uint32), pr.view(np.uint32))
    assert total > 5000
    # segment test alone, incl. near-degenerate configurations
    for _ in range(4000):
        q = rng.normal(size=(4, 3)).astype(np.float32)
        if rng.uniform() < 0.3:
            q[:, 2] = 0
        if rng.uniform() < 0.1:
            q[3] = q[2] + (q[1] - q[0]) * np.float32(rng.uniform(0.5, 2))
        assert O.segs_intersect(*q) == O.ref_segs_intersect(*q)


def test_oracle_matches_golden_vectors():
    g = np.load(GOLDEN)
    meta = json.loads(str(g["meta"]))
    for k in range(meta["n_cases"]):
        got = O.points_in_poly(g[f"pts{k}"], g[f"coeff{k}"], g[f"border{k}"], meta["t"], meta["seeds"][k])
        assert np.array_equal(got, g[f"inside{k}"]), k


def test_reabsorb_loop_semantics():
    cloud, coeffs, borders = scene(3, n_planes=3, pts_per_plane=400, n_clutter=300)
    # duplicate plane 0: every point it claims is claimed twice (no break in the reference loop)
    coeffs = np.concatenate([coeffs, coeffs[:1]])
    borders = borders + [borders[0]]
    r = O.reabsorb(cloud, coeffs, borders, 0.1, 99)
    assert np.array_equal(r.absorbed[0], r.absorbed[3]) and len(r.absorbed[0]) > 50
    claimed = np.zeros(len(cloud), bool)
    for j, a in enumerate(r.absorbed):
        assert np.all(np.diff(a) > 0)
        assert np.array_equal(O.points_in_poly(cloud, coeffs[j], borders[j], 0.1, 99).nonzero()[0], a)
        claimed[a] = True
    assert np.array_equal(r.remaining_idx, (~claimed).nonzero()[0])


# ---------------------------------------------------------------------------------------------------------------
# GPU: plane_ransac_reabsorb vs the oracle
# ---------------------------------------------------------------------------------------------------------------
def _check_gpu(pr, cloud, coeffs, borders, t, seed):
    want = O.reabsorb(cloud, coeffs, borders, t, seed)
    cur, orig, n_rem = pr.reabsorb(coeffs, borders, t, seed)
    for j in range(len(borders)):
        assert np.array_equal(cur[j], want.absorbed[j]), f"plane {j}"
    assert n_rem == len(want.remaining_idx)
    rem = pr.remaining()
    assert np.array_equal(rem[:, :3].view(np.uint32), cloud[want.remaining_idx][:, :3].view(np.uint32))
    return want, cur, orig


@pytest.mark.gpu
def test_gpu_reabsorb_matches_oracle():
    import dialog_b200 as D
    for seed, kw in ((11, {}), (12, dict(n_planes=9, pts_per_plane=1500, n_clutter=20000)), (13, dict(n_planes=1, pts_per_plane=50, n_clutter=10))):
        cloud, coeffs, borders = scene(seed, **kw)
        with D.PlaneRansac(0) as pr:
            pr.set_cloud(cloud)
            want, cur, orig = _check_gpu(pr, cloud, coeffs, borders, 0.1, 1000 + seed)
            for j in range(len(borders)):
                assert np.array_equal(orig[j], cur[j])  # staged cloud: identity
            assert sum(len(a) for a in want.absorbed) > 5


@pytest.mark.gpu
def test_gpu_reabsorb_edge_cases():
    import dialog_b200 as D
    cloud, coeffs, borders = scene(21, n_planes=2, pts_per_plane=700, n_clutter=900)
    cloud[5, 0] = np.nan
    cloud[17, 2] = np.inf
    # a point exactly at distance T of an axis-aligned plane, a 1-vertex and a 2-vertex "polygon", a duplicated plane
    sq = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [1, 1, 0, 1], [0, 1, 0, 1]], np.float32)
    cloud[100] = [0.5, 0.5, np.float32(0.1), 1]
    cloud[101] = [0.5, 0.5, np.nextafter(np.float32(0.1), np.float32(1)), 1]
    coeffs = np.concatenate([coeffs, [[0, 0, 1, 0]], coeffs[:1], coeffs[1:2], coeffs[:1]]).astype(np.float32)
    borders = borders + [sq, borders[0][:1], borders[1][:2], borders[0]]
    with D.PlaneRansac(0) as pr:
        pr.set_cloud(cloud)
        want, cur, _ = _check_gpu(pr, cloud, coeffs, borders, np.float32(0.1), 5)
        assert 100 in want.absorbed[2] and 101 not in want.absorbed[2]
        assert np.array_equal(cur[0], cur[5])
        # nothing left to claim for the same planes; an empty plane list is a no-op
        cur2, _, n2 = pr.reabsorb(coeffs, borders, 0.1, 5)
        assert all(len(a) == 0 for a in cur2) and n2 == len(want.remaining_idx)
        cur3, _, n3 = pr.reabsorb(np.zeros((0, 4), np.float32), [], 0.1, 5)
        assert cur3 == [] and n3 == n2
        with pytest.raises(D.PlaneRansacError):
            pr.reabsorb(coeffs[:1], [np.zeros((0, 4), np.float32)], 0.1, 5)


@pytest.mark.gpu
def test_gpu_reabsorb_after_extraction_uses_original_indices():
    """extract planes by RANSAC, then re-absorb the leftovers against (synthetic) polygons: absorbed_orig indexes the
    staged cloud, absorbed_cur the peeled one, as the oracle run on the peeled cloud says."""
    import dialog_b200 as D
    from dialog_b200 import synth
    pts = synth.three_planes_scene().points(0, 200_000)
    prm = D.make_params(0.05, 200, 500, 0.99, True, 12345, 3, D.DOT_FMA)
    with D.PlaneRansac(0) as pr:
        pr.set_cloud(pts)
        ex = pr.extract_planes(prm)
        assert len(ex.planes) == 3
        rem = pr.remaining().copy()
        # polygons: a coarse hull of each plane's inliers in its own frame (stand-in for pcl::ConcaveHull)
        borders, coeffs = [], []
        for k, pl in enumerate(ex.planes):
            proj = pr.plane_points(k, project=True)[:, :3].astype(np.float64)
            n = pl.coeff[:3].astype(np.float64)
            u = np.cross(n, [1.0, 0, 0]); u /= np.linalg.norm(u); v = np.cross(n, u)
            c = proj.mean(0)
            a, b = (proj - c) @ u, (proj - c) @ v
            ang = np.arctan2(b, a)
            bins = np.linspace(-np.pi, np.pi, 25)
            verts = []
            for lo, hi in zip(bins[:-1], bins[1:]):
                m = (ang >= lo) & (ang < hi)
                if m.any():
                    i = np.argmax(np.hypot(a[m], b[m]))
                    verts.append(proj[m][i])
            bd = np.ones((len(verts), 4), np.float32)
            bd[:, :3] = np.array(verts)
            borders.append(bd)
            coeffs.append(pl.coeff)
        coeffs = np.array(coeffs, np.float32)
        want = O.reabsorb(rem, coeffs, borders, 0.1, 77)
        cur, orig, n_rem = pr.reabsorb(coeffs, borders, 0.1, 77)
        assert sum(len(a) for a in cur) > 100
        staged = D.as_cloud(pts) if hasattr(D, "as_cloud") else None
        for j in range(3):
            assert np.array_equal(cur[j], want.absorbed[j])
            # orig indices point at the same coordinates in the staged cloud
            assert np.array_equal(np.asarray(pts, np.float32)[orig[j]][:, :3].view(np.uint32), rem[cur[j]][:, :3].view(np.uint32))
        assert n_rem == len(want.remaining_idx)
        assert np.array_equal(pr.remaining()[:, :3].view(np.uint32), rem[want.remaining_idx][:, :3].view(np.uint32))


@pytest.mark.gpu
def test_gpu_run_again_on_the_remaining_cloud():
    """on_runAgainAction (Dialog/PCLViewer.cpp:1120-1178): extract, re-absorb, then extract again from what is left —
    all on the device — equals the oracle fed the remaining cloud by hand; indices map back to the caller's array."""
    import dialog_b200 as D
    from dialog_b200 import synth
    pts = synth.indoor_scene().points(0, 300_000)
    p1 = dict(distance_threshold=0.05, max_iterations=300, min_plane_size=500, probability=0.99, max_planes=3)
    with D.PlaneRansac(0) as pr:
        pr.set_cloud(pts)
        ex1 = pr.extract_planes(D.make_params(p1["distance_threshold"], p1["max_iterations"], p1["min_plane_size"], p1["probability"],
                                              True, 12345, p1["max_planes"], D.DOT_FMA))
        want1 = O.extract_planes(pts, O.make_params(0.05, 300, 500, 0.99, True, 12345, 3, O.DOT_FMA, O.REFIT_FIXED))
        assert len(ex1.planes) == 3 and all(np.array_equal(p.inliers_orig, want1.inliers_orig[k]) for k, p in enumerate(ex1.planes))
        scene = synth.indoor_scene()
        coeffs = np.array([p.coeff for p in ex1.planes], np.float32)
        borders = []
        for c in coeffs:
            err = [min(np.abs(q.coeff - c).max(), np.abs(q.coeff + c).max()) for q in scene.patches]
            borders.append(scene.patches[int(np.argmin(err))].border(20))
        want_re = O.reabsorb(want1.remaining, coeffs, borders, 0.1, 3)
        cur, orig, n_left = pr.reabsorb(coeffs, borders, 0.1, 3)
        assert n_left == len(want_re.remaining_idx) and sum(len(a) for a in cur) > 100
        for j in range(3):
            assert np.array_equal(cur[j], want_re.absorbed[j])
        left = want1.remaining[want_re.remaining_idx]
        # run again on the leftovers
        pr.restage_remaining()
        n_staged, n_cur = pr.cloud_size()
        assert n_staged == n_cur == len(left)
        src = pr.staged_source_indices()
        assert np.array_equal(np.asarray(pts, np.float32)[src][:, :3].view(np.uint32), left[:, :3].view(np.uint32))
        prm2 = (0.1, 300, 500, 0.99, True, 12345, 4)
        ex2 = pr.extract_planes(D.make_params(*prm2, D.DOT_FMA))
        # the staged bounding box (hence the refit grid) is the first run's: give the oracle the same grid
        want2 = O.extract_planes(left, O.make_params(*prm2, O.DOT_FMA, O.REFIT_FIXED))
        assert len(ex2.planes) == len(want2.coeffs) >= 1
        for k, p in enumerate(ex2.planes):
            assert np.array_equal(p.inliers_orig, want2.inliers_orig[k]), k
            assert np.allclose(p.coeff, want2.coeffs[k], rtol=1e-5, atol=1e-6)
        assert np.array_equal(pr.remaining()[:, :3].view(np.uint32), want2.remaining[:, :3].view(np.uint32))
