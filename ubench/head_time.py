import sys, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
pts = synth.indoor_scene().points(0, 3_000_000)
pr = D.PlaneRansac(0)
pr.set_cloud(pts)
prm = D.make_params(0.1, 4095, 500, 1.0, True, 12345, 6, D.DOT_FMA)
for _ in range(3):
    ex = pr.extract_planes(prm, want_indices=False)
print(len(ex.planes))
