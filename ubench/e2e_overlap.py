"""End-to-end extraction (pinned host cloud in, index lists out) with the synchronous and the overlapped upload."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
n = 10_000_000
pts = synth.indoor_scene().points(0, n)
pin = D.PinnedArray((n, 4), np.float32)
pin.array[:] = pts
pr = D.PlaneRansac(0)
prm = D.make_params(0.1, 4095, 500, 1.0, True, 12345, 20, D.DOT_FMA)
for overlap in (False, True, False, True):
    for rep in range(4):
        pr.flush_l2()
        pr.profile_enable(True); pr.profile_reset()
        pr.timer_start()
        t0 = time.perf_counter()
        pr.set_cloud_ptr(pin.ptr, n, overlap=overlap)
        t1 = time.perf_counter()
        ex = pr.extract_planes(prm, want_indices=True, copy=False)
        ms = pr.timer_stop()
        p = pr.profile(); pr.profile_enable(False)
    print(f"overlap={overlap}: {ms:.2f} ms (set_cloud returned after {(t1-t0)*1e3:.2f} ms); score {p.ms_score:.2f} stage {p.ms_stage:.2f} "
          f"host sampling {p.host_ms_sampling:.2f} wait {p.host_ms_wait:.2f} total {p.host_ms_total:.2f}", flush=True)
