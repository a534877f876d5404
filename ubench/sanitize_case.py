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
coeff, inl, info = pr.segment_one(D.make_params(0.1, 50, 500, 0.99, True))
print("segment_one", inl.size, info.iterations)
print("plane points", pr.plane_points(0, project=True).shape, "remaining", pr.remaining().shape)
clouds = np.stack([synth.tile_scene(c).points(0, 3000) for c in range(5)])
pr.set_cloud_batch(clouds)
print("batch", pr.segment_batch(D.make_params(0.1, 63, 500, 1.0, True))[1])
pr.close()
print("sanitize case done")
