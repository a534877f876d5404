"""Randomised cross-check of the two peel-loop drivers (device-resident vs host-driven): same planes, lists, remaining cloud."""
import sys, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 150
scenes = [synth.three_planes_scene(), synth.indoor_scene()] + [synth.tile_scene(c) for c in range(3)]
pr = D.PlaneRansac(0)
bad = 0
handed = 0
for case in range(n_cases):
    sc = scenes[int(rng.integers(len(scenes)))]
    n = int(10 ** rng.uniform(3.3, 5.9))
    start = int(rng.integers(0, 1000))
    pts = sc.points(start, start + n).copy()
    if rng.random() < 0.3:
        pts[:: int(rng.integers(5, 40))] = pts[int(rng.integers(n))]      # duplicates: degenerate samples
    if rng.random() < 0.2:
        pts[rng.integers(0, n, 20), int(rng.integers(3))] = np.nan
    K = int(2 ** rng.integers(4, 13))
    prm = D.make_params(float(rng.choice([0.05, 0.1, 0.02])), K - 1, int(rng.choice([50, 500, 3000])), 1.0, bool(rng.random() < 0.8), 12345 if rng.random() < 0.7 else int(rng.integers(1, 1 << 30)),
                        int(rng.integers(1, 9)), int(rng.integers(0, 2)))
    pr.set_cloud(pts)
    a = pr.extract_planes(prm); ra = pr.remaining().copy()
    pr.set_round_loop(host=True)
    b = pr.extract_planes(prm); rb = pr.remaining().copy()
    pr.set_round_loop(host=False)
    same = len(a.planes) == len(b.planes) and ra.tobytes() == rb.tobytes() and len(a.infos) == len(b.infos) and all(
        np.array_equal(p.coeff.view(np.uint32), q.coeff.view(np.uint32)) and np.array_equal(p.inliers_orig, q.inliers_orig) and np.array_equal(p.inliers_cur, q.inliers_cur)
        and list(p.info.best_sample) == list(q.info.best_sample) and p.info.best_count == q.info.best_count and p.info.draws == q.info.draws
        for p, q in zip(a.planes, b.planes))
    handed += any(i.draws > i.iterations for i in a.infos)
    if not same:
        bad += 1
        print("MISMATCH case", case, "n", n, "K", K, "planes", len(a.planes), len(b.planes), flush=True)
print(f"{n_cases} random cases, {bad} mismatches, {handed} with rounds handed back for redraws")
sys.exit(1 if bad else 0)
