"""PCL-default adaptive mode (max_iterations=50, probability=0.99): where a 10M-point, 20-plane extraction spends its time."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
pts = synth.indoor_scene().points(0, 10_000_000)
pr = D.PlaneRansac(0)
pr.set_cloud(pts)
prm = D.make_params(0.1, 50, 500, 0.99, True, 12345, 20, D.DOT_FMA)
for prof_on in (False, True):
    pr.profile_enable(prof_on)
    pr.extract_planes(prm, want_indices=False)
    pr.profile_reset()
    t0 = time.perf_counter(); ex = pr.extract_planes(prm, want_indices=False); dt = time.perf_counter() - t0
    p = pr.profile()
    print(f"profiling={prof_on} wall {dt*1e3:.2f} ms; host total {p.host_ms_total:.2f} sampling {p.host_ms_sampling:.2f} replay {p.host_ms_replay:.2f} wait {p.host_ms_wait:.2f}; "
          f"kernels: score {p.ms_score:.2f} models {p.ms_models:.2f} refit {p.ms_refit:.2f} compact {p.ms_compact:.2f}; launches score {p.launches_score} scored per round {[int(i.n_scored) for i in ex.infos][:5]}")
