"""configs[4] slice: 512 clouds of 32K points, 256 hypotheses each, one GPU; whole call without events, then by kernel class."""
import sys, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
nc, n_per, K = 512, 32768, 256
pin = D.PinnedArray((nc, n_per, 4), np.float32)
for j in range(nc):
    pin.array[j] = synth.tile_scene(j).points(0, n_per)
pr = D.PlaneRansac(0)
peak = pr.measure_ffma_peak()
pr.set_cloud_batch_ptr(pin.ptr, nc, n_per)
prm = D.make_params(0.1, K - 1, 500, 1.0, True, 12345, 1, D.DOT_FMA)
for _ in range(3):
    pr.segment_batch(prm, want_infos=False)
ms = []
for _ in range(8):
    pr.flush_l2(); pr.timer_start(); pr.segment_batch(prm, want_infos=False); ms.append(pr.timer_stop())
pr.profile_enable(True); pr.profile_reset()
for _ in range(5):
    pr.flush_l2(); pr.segment_batch(prm, want_infos=False)
p = pr.profile()
step = float(np.median(ms))
print(f"{step:.4f} ms per batch = {100 * 6 * nc * n_per * K / (step * 1e-3) / 1e12 / peak:.1f}% of FP32 peak; events: gather+models {p.ms_models / 5:.4f} score {p.ms_score / 5:.4f} "
      f"refit {p.ms_refit / 5:.4f} count {p.ms_compact / 5:.4f} other {p.ms_other / 5:.4f}")
