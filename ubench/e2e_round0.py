"""Where the end-to-end overhead of the first round goes: one-plane extraction, resident against uploaded."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
n = 10_000_000
pts = synth.indoor_scene().points(0, n)
pin = D.PinnedArray((n, 4), np.float32)
pin.array[:] = pts
pr = D.PlaneRansac(0)
for planes in (1, 2):
    prm = D.make_params(0.1, 4095, 500, 1.0, True, 12345, planes, D.DOT_FMA)
    for mode in ("resident-device-loop", "resident-host-loop", "uploaded"):
        pr.set_round_loop(mode == "resident-host-loop")
        ms = []
        for rep in range(7):
            if mode != "uploaded":
                pr.set_cloud_ptr(pin.ptr, n)
            pr.flush_l2()
            pr.profile_enable(rep == 6); pr.profile_reset()
            t0 = time.perf_counter()
            pr.timer_start()
            if mode == "uploaded":
                pr.set_cloud_ptr(pin.ptr, n, overlap=True)
            t1 = time.perf_counter()
            ex = pr.extract_planes(prm, want_indices=False)
            t2 = time.perf_counter()
            v = pr.timer_stop()
            if rep < 6:
                ms.append(v)
            p = pr.profile()
        print(f"planes={planes} {mode}: {np.median(ms[2:]):.3f} ms; with events: score {p.ms_score:.3f} stage {p.ms_stage:.3f} models {p.ms_models:.3f} "
              f"refit {p.ms_refit:.3f} compact {p.ms_compact:.3f} other {p.ms_other:.3f} | host: set_cloud {(t1 - t0) * 1e3:.3f} extract {(t2 - t1) * 1e3:.3f} "
              f"sampling {p.host_ms_sampling:.3f} replay {p.host_ms_replay:.3f} wait {p.host_ms_wait:.3f} total {p.host_ms_total:.3f}", flush=True)
pr.set_round_loop(False)
