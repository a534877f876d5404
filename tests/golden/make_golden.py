"""Regenerates the fixtures under tests/golden/.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

double_shadow_xyz.npy   the x,y,z columns of the reference's only fixture on this path,
                        Dialog/double_shadow.pcd (ASCII PCD v0.7, FIELDS x y z rgb, 991 points), as
                        float32 — config 1 of BASELINE.json.  Only the coordinates are kept.
double_shadow_golden.json   what the CPU oracle (oracle/pr_oracle.c) returns for that cloud with the
                        reference's parameters (Dialog/config.txt:29 T_dist_point_plane = 0.1) and with the
                        discriminating threshold 0.005, PCL defaults otherwise.  The
                        reference records no outputs for this path, so these pin the oracle against
                        regressions, not against PCL ("parity unpinned", oracle/pr_oracle.h).
synthetic_golden.json   oracle results on a 20 000-point slice of the config-2 scene.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from dialog_b200 import synth  # noqa: E402

REF_PCD = "/root/reference/Dialog/double_shadow.pcd"


def read_ascii_pcd_xyz(path):
    with open(path, "r") as f:
        lines = f.read().splitlines()
    i = next(k for k, l in enumerate(lines) if l.startswith("DATA"))
    assert lines[i].split()[1] == "ascii"
    n = int(next(l for l in lines if l.startswith("POINTS")).split()[1])
    rows = [l.split()[:3] for l in lines[i + 1: i + 1 + n]]
    return np.array(rows, dtype=np.float32)


def hexf(a):
    return [float(v).hex() for v in np.asarray(a, np.float32)]


def run_cases(pts, cases):
    out = []
    for c in cases:
        prm = O.make_params(**c)
        seg = O.segment(pts, prm)
        out.append(dict(params=c, ok=bool(seg.ok), coeff=hexf(seg.coeff), n_inliers=int(seg.inliers.size),
                        inliers_sha=int(np.bitwise_xor.reduce(seg.inliers.astype(np.int64) * 2654435761 % (1 << 32))) if seg.inliers.size else 0,
                        iterations=seg.trace.iterations, draws=seg.trace.draws, best_sample=list(seg.trace.best_sample),
                        best_count=seg.trace.best_count, raw_coeff=hexf(list(seg.trace.raw_coeff))))
    return out


def main():
    xyz = read_ascii_pcd_xyz(REF_PCD)
    assert xyz.shape == (991, 3)
    np.save(os.path.join(HERE, "double_shadow_xyz.npy"), xyz)
    cases = []
    for t in (0.1, 0.005):
        for dot in (O.DOT_PCL_SSE2, O.DOT_FMA):
            for refit in (O.REFIT_PCL_FLOAT, O.REFIT_FIXED):
                cases.append(dict(distance_threshold=t, max_iterations=50, min_plane_size=500, probability=0.99,
                                  optimize_coefficients=True, seed=12345, max_planes=8, dot_order=dot, refit_mode=refit))
    json.dump(run_cases(xyz, cases), open(os.path.join(HERE, "double_shadow_golden.json"), "w"), indent=1)

    pts = synth.three_planes_scene().points(0, 20000)
    cases = []
    for dot in (O.DOT_PCL_SSE2, O.DOT_FMA):
        for prob, it in ((0.99, 50), (1.0, 255)):
            cases.append(dict(distance_threshold=0.1, max_iterations=it, min_plane_size=500, probability=prob,
                              optimize_coefficients=True, seed=12345, max_planes=8, dot_order=dot,
                              refit_mode=O.REFIT_FIXED))
    json.dump(run_cases(pts, cases), open(os.path.join(HERE, "synthetic_golden.json"), "w"), indent=1)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
