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
    # the device-resident loop without events: the kernels' own %globaltimer stamps (pr_profile.loop_ms)
    pr.set_round_loop(False)
    pr.profile_enable(False)
    for _ in range(3): pr.extract_planes(prm, want_indices=False)
    pr.profile_reset()
    ms = []
    for _ in range(5):
        pr.flush_l2(); pr.timer_start(); ex = pr.extract_planes(prm, want_indices=False); ms.append(pr.timer_stop())
    p = pr.profile()
    step = sum(ms) / 5
    pairs = sum(int(i.n_cloud) * int(i.n_scored) for i in ex.infos)
    stages = ", ".join(f"{nm} {v / 5:.4f}" for nm, v in zip(D.LOOP_STAGE_NAMES, p.loop_ms) if v > 0)
    print(f"{name} loop=device, no events: {step:.3f} ms/extraction = {100 * 6 * pairs / (step * 1e-3) / 1e12 / peak:.1f}% of FP32 peak end to end; "
          f"{p.loop_rounds / 5:.0f} rounds, device stages (ms): {stages}; rounds total {sum(p.loop_ms) / 5:.3f}, step - rounds {step - sum(p.loop_ms) / 5:.3f}")
    tl = pr.round_timeline().astype(np.int64)
    for r, row in enumerate(tl[:4]):
        used = [(nm, v) for nm, v in zip(list(D.LOOP_STAGE_NAMES) + ["end"], row) if v > 0]
        print("   round", r, " ".join(f"{a[0]}={(b[1] - a[1]) / 1e3:.1f}us" for a, b in zip(used, used[1:])))
    pr.close()
