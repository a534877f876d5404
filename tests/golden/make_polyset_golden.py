"""Golden vectors on the reference's OWN polygons: Dialog/dataForPlane/source_plane_registration.{pcd,txt} holds the 11
plane borders (1960 vertices) its detector saved for the registration path.  This script stores them compactly and
records, for seeded points around every polygon, what the reference's isPointInPoly source (oracle/build_ref.py) says.
Run in the build container (needs /root/reference):  python tests/golden/make_polyset_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dialog_b200 import polyset  # noqa: E402
from oracle import oracle as O  # noqa: E402

SRC = "/root/reference/Dialog/dataForPlane/source_plane_registration.pcd"
if O.ref_lib() is None or not os.path.exists(SRC):
    raise SystemExit("needs /root/reference")
borders, normals, scales = polyset.load_polygon_set(SRC)
assert normals is None and len(borders) == 11 and sum(len(b) for b in borders) == 1960
rng = np.random.default_rng(20261019)
out = {"vertices": np.concatenate(borders)[:, :3].astype(np.float32), "sizes": np.array([len(b) for b in borders], np.int32)}
seeds, thresholds = [], []
for k, b in enumerate(borders):
    coeff = polyset.plane_through_border(b)
    v = b[:, :3].astype(np.float64)
    c, ext = v.mean(0), np.ptp(v, axis=0).max()
    n = coeff[:3].astype(np.float64)
    u = np.cross(n, [1.0, 0, 0]); u /= np.linalg.norm(u); w = np.cross(n, u)
    m = 600
    ab = rng.uniform(-0.7 * ext, 0.7 * ext, (m, 2))
    t = np.float32(0.02 * ext)   # the data set is in its own units: scale the distance threshold with the polygon
    h = rng.normal(0, 0.8 * t, m)
    pts = np.ones((m, 4), np.float32)
    pts[:, :3] = c + np.outer(ab[:, 0], u) + np.outer(ab[:, 1], w) + np.outer(h, n)
    seed = int(rng.integers(0, 2**32))
    out[f"coeff{k}"], out[f"pts{k}"] = coeff, pts
    out[f"inside{k}"] = O.ref_points_in_poly(pts, coeff, b, t, seed)
    seeds.append(seed)
    thresholds.append(float(t))
    print(k, len(b), float(ext), int(out[f"inside{k}"].sum()))
out["meta"] = json.dumps({"seeds": seeds, "thresholds": thresholds,
                          "source": "Dialog/dataForPlane/source_plane_registration.{pcd,txt}; inside = reference isPointInPoly (oracle/_ref)"})
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_polygons_golden.npz"), **out)
