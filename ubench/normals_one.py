import sys, numpy as np
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
pr = D.PlaneRansac(0)
pts = synth.indoor_scene().points(0, 2_000_000)
pr.set_cloud(pts)
nrm, cnt = pr.estimate_normals(0.25, want_counts=True)
print("mean neighbours", cnt.mean())
r, l = pr.cluster_filter(0.1, 500)
print("cluster filter removed", r, "left", l)
