import sys, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
base = synth.indoor_scene().points(0, 10_000_000)
pr = D.PlaneRansac(0)
hbm = pr.measure_copy_bw(1 << 30)
print(f"copy kernel bandwidth (read+write): {hbm:.0f} GB/s")
for rep in (1, 5, 10):
    pts = np.tile(base, (rep, 1)) if rep > 1 else base
    n = pts.shape[0]
    pr.set_cloud(pts)
    prm = D.make_params(0.1, 255, 500, 1.0, True, 12345, 1, D.DOT_FMA)
    pr.extract_planes(prm, want_indices=False)
    pr.profile_enable(True); pr.profile_reset()
    for _ in range(5): ex = pr.extract_planes(prm, want_indices=False)
    p = pr.profile(); pr.profile_enable(False)
    print(f"N={n}: compact {p.ms_compact/5*1e3:.1f} us {p.bytes_compact/(p.ms_compact*1e-3)/1e9:.0f} GB/s | refit {p.ms_refit/5*1e3:.1f} us {p.bytes_refit/(p.ms_refit*1e-3)/1e9:.0f} GB/s | inliers {ex.planes[0].info.n_inliers}")
