import sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
import dialog_b200 as D
from dialog_b200 import synth
sc = synth.three_planes_scene()
pr = D.PlaneRansac(0)
for n, K in [(100000, 256), (100000, 255), (100000, 128), (100000, 129), (300001, 1024), (50000, 2500)]:
    pts = sc.points(0, n)
    tri = O.draw_sequence(n, K)
    pr.set_cloud(pts)
    oc, og = O.models_from_triples(pts, tri)
    for order in (0, 1):
        counts = pr.score(tri, 0.1, order)
        want = O.count_batch(pts, np.nan_to_num(oc), 0.1, order, threads=8)
        bad = np.nonzero(counts != want)[0]
        print(n, K, order, "mismatches", bad.size, bad[:10], (counts - want)[bad[:10]])
