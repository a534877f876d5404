import sys, os, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
pts = synth.indoor_scene().points(0, n)
pr = D.PlaneRansac(0)
pr.set_cloud(pts)
for name, sc in (("brute", D.SCORER_BRUTE), ("hier", D.SCORER_HIER)):
    prm = D.make_params(0.1, K - 1, 500, 1.0, False, 12345, 1, D.DOT_FMA, sc)
    pr.segment_one(prm); pr.segment_one(prm)
    pr.profile_enable(True); pr.profile_reset()
    for _ in range(3): coeff, inl, info = pr.segment_one(prm)
    p = pr.profile(); pr.profile_enable(False)
    print(f"{name}: N={n} K={K}: score {p.ms_score/3:.3f} ms/call  ({p.pairs_scored/(p.ms_score*1e-3)/1e12:.2f} Tpairs/s)  best_count {info.best_count}")
