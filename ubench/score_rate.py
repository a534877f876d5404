import sys, os, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pts = synth.indoor_scene().points(0, n)
pr = D.PlaneRansac(0)
pr.set_cloud(pts)
peak = pr.measure_ffma_peak()
for K in (256, 1024, 4096):
    tri = D.host_draw_triples(n, K)
    pr.score(tri, 0.1); pr.score(tri, 0.1)
    pr.profile_enable(True); pr.profile_reset()
    for _ in range(5): c = pr.score(tri, 0.1)
    p = pr.profile(); pr.profile_enable(False)
    tf = 6.0 * p.pairs_scored / (p.ms_score * 1e-3) / 1e12
    print(f"PR_SCORE_H={os.environ.get('PR_SCORE_H','-')} N={n} K={K}: {p.ms_score/5:.3f} ms/launch  {tf:.2f} TF  {100*tf/peak:.1f}% of {peak:.1f}  checksum {int(c.sum())}")
