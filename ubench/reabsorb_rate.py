"""Re-absorption pass (postProcessPlanes) alone: 10M-point indoor scene, extraction at t=0.05, re-absorption at T=0.1."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene = synth.indoor_scene()
pts = scene.points(0, n)
pr = D.PlaneRansac(0)
pr.set_cloud(pts)
ext = pr.extract_planes(D.make_params(0.05, 50, 500, 0.99, True, 12345, 20, D.DOT_FMA), want_indices=False)
rem = pr.remaining().copy()
coeffs = np.array([p.coeff for p in ext.planes], np.float32)
borders = []
for c in coeffs:
    err = [min(np.abs(q.coeff - c).max(), np.abs(q.coeff + c).max()) for q in scene.patches]
    borders.append(scene.patches[int(np.argmin(err))].border())
for rep in range(reps):
    pr.set_cloud(rem)
    pr.profile_enable(True); pr.profile_reset()
    pr.timer_start()
    cur, _, n_left = pr.reabsorb(coeffs, borders, 0.1, 20261018)
    ms = pr.timer_stop()
    p = pr.profile(); pr.profile_enable(False)
    print(f"N_rem={len(rem)} planes={len(coeffs)} absorbed={sum(len(a) for a in cur)} left={n_left}: {ms:.3f} ms  (kernels: other {p.ms_other:.3f} compact {p.ms_compact:.3f})", flush=True)
