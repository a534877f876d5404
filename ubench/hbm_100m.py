import sys, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
# one segment() + peel over 100M points (ten copies of the 10M-point storey), for the ncu capture of K3 / K5 at that size
base = synth.indoor_scene().points(0, 10_000_000)
pts = np.tile(base, (10, 1))
pr = D.PlaneRansac(0)
pr.set_cloud(pts)
prm = D.make_params(0.1, 63, 500, 1.0, True, 12345, 1, D.DOT_FMA)
pr.extract_planes(prm, want_indices=False)
pr.profile_enable(True); pr.profile_reset()
for _ in range(2): ex = pr.extract_planes(prm, want_indices=False)
p = pr.profile()
print(f"N={pts.shape[0]}: compact {p.ms_compact/2*1e3:.1f} us, {p.bytes_compact//2} B, {p.bytes_compact/(p.ms_compact*1e-3)/1e9:.0f} GB/s | "
      f"refit {p.ms_refit/2*1e3:.1f} us, {p.bytes_refit//2} B, {p.bytes_refit/(p.ms_refit*1e-3)/1e9:.0f} GB/s | inliers {ex.planes[0].info.n_inliers}")
