"""End-to-end extraction (pinned host cloud in, index lists out) against the resident one, for one upload chunk schedule
(PR_UPLOAD_WEIGHTS, read once per process).  No event profiler."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
n = 10_000_000
pts = synth.indoor_scene().points(0, n)
pin = D.PinnedArray((n, 4), np.float32)
pin.array[:] = pts
pr = D.PlaneRansac(0)
prm = D.make_params(0.1, 4095, 500, 1.0, True, 12345, 20, D.DOT_FMA)
pr.set_cloud_ptr(pin.ptr, n)
res = []
for rep in range(6):
    pr.flush_l2(); pr.timer_start(); pr.extract_planes(prm, want_indices=False); res.append(pr.timer_stop())
e2e = []
for rep in range(8):
    pr.flush_l2()
    pr.timer_start()
    pr.set_cloud_ptr(pin.ptr, n, overlap=True)
    ex = pr.extract_planes(prm, want_indices=True, copy=False)
    e2e.append(pr.timer_stop())
nol = []
for rep in range(6):
    pr.flush_l2()
    pr.timer_start()
    pr.set_cloud_ptr(pin.ptr, n, overlap=True)
    pr.extract_planes(prm, want_indices=False)
    nol.append(pr.timer_stop())
print(f"   without the index lists: {np.median(nol[2:]):.3f} ms")
print(f"weights={os.environ.get('PR_UPLOAD_WEIGHTS', 'default')}: resident {np.median(res[2:]):.3f} ms, end to end {np.median(e2e[3:]):.3f} ms "
      f"(min {min(e2e[3:]):.3f}), difference {np.median(e2e[3:]) - np.median(res[2:]):.3f} ms", flush=True)
