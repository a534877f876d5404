import sys, time, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
# host-phase timers against kernel time, for the device-driven (default) and the host-driven round loop
for name, scene, n, K, planes in (("configs[2] 10M x 4096 x 20", synth.indoor_scene(), 10_000_000, 4096, 20),
                                 ("configs[1] 1M x 1024 x 3", synth.three_planes_scene(), 1_000_000, 1024, 3)):
    pts = scene.points(0, n)
    pr = D.PlaneRansac(0)
    pr.set_cloud(pts)
    peak = pr.measure_ffma_peak()
    prm = D.make_params(0.1, K - 1, 500, 1.0, True, 12345, planes, D.DOT_FMA)
    for host_loop in (False, True):
        pr.set_round_loop(host_loop)
        pr.profile_enable(True)
        for _ in range(3): pr.extract_planes(prm, want_indices=False)
        pr.profile_reset()
        ms = []
        for _ in range(5):
            pr.flush_l2(); pr.timer_start(); ex = pr.extract_planes(prm, want_indices=False); ms.append(pr.timer_stop())
        p = pr.profile()
        pairs = sum(int(i.n_cloud) * int(i.n_scored) for i in ex.infos)
        step = sum(ms) / 5
        kern = (p.ms_score + p.ms_models + p.ms_refit + p.ms_compact + p.ms_other) / 5
        print(f"{name} loop={'host' if host_loop else 'device'}: {step:.3f} ms/extraction = {100 * 6 * pairs / (step * 1e-3) / 1e12 / peak:.1f}% of FP32 peak "
              f"end to end; kernels {kern:.3f} ms (score {p.ms_score/5:.3f} = {100 * 6 * p.pairs_scored / (p.ms_score * 1e-3) / 1e12 / peak:.1f}%, models+draw {p.ms_models/5:.3f}, "
              f"refit {p.ms_refit/5:.3f}, compact {p.ms_compact/5:.3f}, other {p.ms_other/5:.3f}); gap {step - kern:.3f} ms; host wait {p.host_ms_wait/5:.2f} sampling {p.host_ms_sampling/5:.2f}")
    pr.close()
