"""CPU tests of the host-side logic the library exports (no device needed), checked against the oracle:
PCL's sampling stream, the sequential RANSAC decisions replayed over batched counts, the plane from
integer moments, shard ranges."""
import numpy as np
import pytest

import dialog_b200 as D


@pytest.fixture(autouse=True)
def _lib(lib_built):
    return lib_built


@pytest.mark.parametrize("n", [3, 4, 10, 991, 65536, 10_000_000])
def test_sampler_matches_oracle(O, n):
    draws = 300
    assert (D.host_draw_triples(n, draws) == O.draw_sequence(n, draws)).all()
    assert (D.host_draw_triples(n, 20, seed=7) == O.draw_sequence(n, 20, seed=7)).all()


@pytest.mark.parametrize("n,draws", [(3, 5), (3, 400), (4, 100), (10, 400), (100, 300), (1000, 600), (32768, 256), (32768, 640),
                                     (100_000, 4096), (1_000_000, 4096), (2_200_000, 4096), (10_000_000, 4096),
                                     (2**31 - 5000, 4096)])
def test_parallel_sampler_equals_the_sequential_walk(n, draws):
    """The formulation the device-side round loop runs (csrc/pr_draw.h, same per-op code emulated on the host): independent
    ops + a sequential replay of the colliding ones gives PCL's triples exactly, down to clouds where every op collides."""
    for seed in (12345, 7):
        got = D.host_draw_triples_parallel(n, draws, seed=seed)
        assert got is not None
        assert (got == D.host_draw_triples(n, draws, seed=seed)).all()


def test_parallel_sampler_hands_crowded_rounds_back():
    # ~ (3K)^2 / N colliding ops: more than the device replays sequentially -> the round goes to the host sampler
    assert D.host_draw_triples_parallel(5000, 4096) is None
    rng = np.random.default_rng(3)
    for _ in range(60):
        n, k = int(10 ** rng.uniform(0.5, 7)), int(rng.integers(1, 3000))
        got = D.host_draw_triples_parallel(n, k)
        assert got is None or (got == D.host_draw_triples(n, k)).all()


def test_parallel_sampler_dependency_chains_random_seeds():
    """Rounds the size the bench runs (tens to hundreds of colliding picks among 3K ops, the listed ops often next to each
    other in time): without a swap inside the head every listed op is resolved by following its own dependency chain; with
    one, by the sequential replay.  Both must give PCL's triples for every seed."""
    rng = np.random.default_rng(11)
    for n, k, reps in ((1_000_000, 4096, 40), (300_000, 2048, 40), (2_000_000, 8192, 15), (100_000, 1000, 60), (50_000, 512, 60)):
        for _ in range(reps):
            seed = int(rng.integers(1, 2**31 - 1))
            got = D.host_draw_triples_parallel(n, k, seed=seed)
            assert got is not None and (got == D.host_draw_triples(n, k, seed=seed)).all(), (n, k, seed)


def test_sampler_rejects_tiny_clouds():
    with pytest.raises(D.PlaneRansacError):
        D.host_draw_triples(2, 1)


def _oracle_counts(O, pts, n_draws, t, order):
    tri = O.draw_sequence(pts.shape[0], n_draws)
    coeffs, good = O.models_from_triples(pts, tri)
    counts = O.count_batch(pts, np.nan_to_num(coeffs, nan=0.0), t, order)
    counts[~good] = 0
    return tri, coeffs, good, counts


@pytest.mark.parametrize("prob,max_it", [(0.99, 50), (0.99, 1000), (1.0, 199), (0.5, 10), (0.99, 0)])
def test_replay_reproduces_compute_model(O, scene2, prob, max_it):
    pts = scene2.points(0, 20_000)
    prm = O.make_params(0.1, max_it, 500, prob, False, 12345, 8, O.DOT_FMA, O.REFIT_FIXED)
    seg = O.segment(pts, prm)
    tri, coeffs, good, counts = _oracle_counts(O, pts, max_it + 1 + 8, 0.1, O.DOT_FMA)
    r = D.host_replay(counts, good, pts.shape[0], max_it, prob)
    assert not r["exhausted"]
    assert r["iterations"] == seg.trace.iterations and r["draws_used"] == seg.trace.draws
    if seg.ok:
        assert tri[r["best_draw"]].tolist() == list(seg.trace.best_sample)
        assert counts[r["best_draw"]] == seg.trace.best_count
    else:
        assert r["best_draw"] == -1


def test_replay_skips_bad_draws_and_gives_up_after_1000(O):
    counts = np.array([5, 0, 9, 9, 3], np.int32)
    good = np.array([1, 0, 1, 1, 1], np.uint8)
    r = D.host_replay(counts, good, 100, 3, 1.0)       # 4 trials = 5 draws, one of them rejected
    assert (r["best_draw"], r["iterations"], r["draws_used"], r["exhausted"]) == (2, 4, 5, False)   # strict '>': first 9 wins
    r = D.host_replay(counts, good, 100, 10, 1.0)
    assert r["exhausted"] and r["iterations"] == 4
    bad = np.zeros(1500, np.uint8)
    r = D.host_replay(np.zeros(1500, np.int32), bad, 100, 50, 0.99)
    assert (r["best_draw"], r["draws_used"], r["exhausted"]) == (-1, 1000, False)


def test_plane_from_moments_matches_oracle_bitwise(O, scene2):
    pts = scene2.points(0, 50_000)
    s = O.fixed_scale_exp(pts)
    for j, patch in enumerate(scene2.patches):
        c0 = patch.coeff.astype(np.float32)
        idx = O.select_within(pts, c0, 0.1, O.DOT_FMA)
        piv = pts[idx[0], :3]
        want, mom = O.refit_fixed(pts, idx, piv, s, c0)
        got = D.host_plane_from_moments(mom, piv, s)
        assert got.tobytes() == want.tobytes()
        # any (hi, lo) split of the same totals gives the same plane (the device produces a different split)
        m2 = mom.copy()
        for k in range(6):
            m2[4 + 2 * k] -= 3
            m2[5 + 2 * k] += 3 << 32
        assert D.host_plane_from_moments(m2, piv, s).tobytes() == want.tobytes()


def test_pcl_float_solve_matches_oracle_bitwise(O, scene2, lib_built):
    """PR_REFIT_PCL_FLOAT's host half (FP32 eigen33 of the nine sequential sums) against the oracle's ORC_REFIT_PCL_FLOAT;
    the sums themselves are formed here the way the device thread forms them: one FP32 addition at a time, in index order."""
    import ctypes as C
    from dialog_b200 import _lib
    L = _lib.load()
    pts = scene2.points(0, 60_000)
    for patch in scene2.patches:
        c0 = patch.coeff.astype(np.float32)
        idx = O.select_within(pts, c0, 0.1, O.DOT_FMA)
        x, y, z = (pts[idx, a] for a in range(3))
        terms = [x * x, x * y, x * z, y * y, y * z, z * z, x, y, z]            # float32 products
        sums = np.array([np.cumsum(t, dtype=np.float32)[-1] for t in terms], np.float32)   # cumsum adds sequentially
        got = np.zeros(4, np.float32)
        _lib.check(L.plane_ransac_host_plane_from_pcl_float_sums(sums.ctypes.data_as(C.c_void_p), idx.size, got.ctypes.data_as(C.c_void_p)))
        assert got.tobytes() == O.refit_pcl_float(pts, idx, c0).tobytes()


def test_plane_from_moments_needs_four_points():
    with pytest.raises(D.PlaneRansacError):
        D.host_plane_from_moments(np.zeros(16, np.int64), np.zeros(3, np.float32), 0)


@pytest.mark.parametrize("n,r", [(0, 1), (10, 3), (100_000_000, 8), (7, 8)])
def test_shard_ranges_tile_the_cloud(n, r):
    nxt = 0
    for k in range(r):
        first, count = D.host_shard_range(n, r, k)
        assert first == nxt and count >= 0
        nxt += count
    assert nxt == n
    sizes = [D.host_shard_range(n, r, k)[1] for k in range(r)]
    assert max(sizes) - min(sizes) <= 1


def test_host_rand_edges_is_the_msvc_generator(lib_built):
    """isPointInPoly's edge draws (Dialog/PlaneDetect.h:1905-1919): srand(seed), rand() % border.size() with the MSVC CRT
    generator — the library's host helper against the oracle's and against the generator's well-known first values."""
    import dialog_b200 as D
    from oracle import oracle as O
    assert D.host_rand_edges(1, 32768).tolist() == [41, 18467, 6334, 26500, 19169, 15724, 11478, 29358, 26962, 24464]
    for seed in (0, 1, 12345, 2**31 - 1, 2**32 - 1, 1729):
        for nb in (1, 2, 3, 77, 337, 100000):
            assert np.array_equal(D.host_rand_edges(seed, nb), O.msvc_rand_edges(seed, nb)), (seed, nb)
    with pytest.raises(D.PlaneRansacError):
        D.host_rand_edges(1, 0)
