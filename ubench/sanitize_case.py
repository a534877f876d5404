"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel once, tiny sizes."""
import sys
import numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth

pts = synth.three_planes_scene().points(0, 20_011)
pts[::501, 0] = np.nan
pr = D.PlaneRansac(0)
kept, cen, src = pr.set_cloud_preprocessed(pts)
for sc in (D.SCORER_BRUTE, D.SCORER_HIER):
    for order in (D.DOT_FMA, D.DOT_PCL_SSE2):
        ex = pr.extract_planes(D.make_params(0.1, 300, 1000, 1.0, True, 12345, 4, order, sc))
        print("scorer", sc, "order", order, "planes", len(ex.planes), [p.inliers_orig.size for p in ex.planes])
# both round loops (device-resident by default; host-driven), the PCL-float refit, a crowded round that is handed back
pr.set_round_loop(host=True)
ex_h = pr.extract_planes(D.make_params(0.1, 300, 1000, 1.0, True, 12345, 4))
pr.set_round_loop(host=False)
print("host loop planes", len(ex_h.planes), "same", all(a.coeff.tobytes() == b.coeff.tobytes() for a, b in zip(ex_h.planes, ex.planes)))
ex_p = pr.extract_planes(D.make_params(0.1, 300, 1000, 1.0, True, 12345, 4, D.DOT_FMA, D.SCORER_BRUTE, D.REFIT_PCL_FLOAT))
print("pcl-float refit planes", len(ex_p.planes))
ex_c = pr.extract_planes(D.make_params(0.1, 1023, 1000, 1.0, True, 12345, 2))
print("K=1024 on 20k points (crowded sampler)", len(ex_c.planes))
coeff, inl, info = pr.segment_one(D.make_params(0.1, 50, 500, 0.99, True))
print("segment_one", inl.size, info.iterations)
print("plane points", pr.plane_points(0, project=True).shape, "remaining", pr.remaining().shape)
# re-absorption pass + run again
coeffs = np.array([p.coeff for p in ex.planes], np.float32)
borders = []
for c in coeffs:
    n = c[:3].astype(np.float64); u = np.cross(n, [1.0, 0, 0]); u /= np.linalg.norm(u); v = np.cross(n, u)
    o = np.array([1.5, 1.5, 1.5]); o = o - (n @ o + c[3]) * n
    borders.append(np.array([o + 1.4 * (np.cos(a) * u + np.sin(a) * v) for a in np.linspace(0, 2 * np.pi, 37, endpoint=False)], np.float32))
cur, orig, left = pr.reabsorb(coeffs, borders, 0.2, 42)
print("reabsorb", [a.size for a in cur], "left", left)
pr.restage_remaining()
ex = pr.extract_planes(D.make_params(0.1, 100, 200, 1.0, True, 12345, 2))
print("run again", len(ex.planes), pr.staged_source_indices().shape)
clouds = np.stack([synth.tile_scene(c).points(0, 3000) for c in range(5)])
pr.set_cloud_batch(clouds)
print("batch", pr.segment_batch(D.make_params(0.1, 63, 500, 1.0, True))[1])
bc, bn, bi, bl = pr.segment_batch(D.make_params(0.1, 63, 500, 1.0, True), want_lists=True)
print("batch lists", [l.size for l in bl])
nrm, cnt = pr.estimate_normals(0.2, want_counts=True) if False else (None, None)
pr.close()
print("sanitize case done")
