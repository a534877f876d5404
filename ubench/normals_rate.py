"""Radius-PCA normal estimation (estimateNormal of the reference) on the indoor scene."""
import sys, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
pr = D.PlaneRansac(0)
for n, radius in ((1_000_000, 0.5), (10_000_000, 0.1), (10_000_000, 0.25)):
    pts = synth.indoor_scene().points(0, n)
    pr.set_cloud(pts)
    nrm, cnt = pr.estimate_normals(radius, want_counts=True)
    ms = []
    for _ in range(2):
        pr.timer_start()
        nrm = pr.estimate_normals(radius)
        ms.append(pr.timer_stop())
    pairs = float(cnt.astype(np.int64).sum())
    print(f"N={n} r={radius}: {min(ms):.1f} ms, mean neighbours {cnt.mean():.0f}, max {cnt.max()}, {pairs/(min(ms)*1e-3):.3e} neighbour pairs/s, "
          f"NaN normals {int(np.isnan(nrm[:,0]).sum())}", flush=True)
