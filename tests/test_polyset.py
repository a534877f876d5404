"""The reference's polygon-set files (Dialog/PCLViewer.cpp:1341-1396, 1005-1100) and the re-absorption
pass on the reference's OWN polygons (Dialog/dataForPlane: 11 plane borders, 1960 vertices, saved by its detector)."""
import json
import os

import numpy as np
import pytest

from dialog_b200 import polyset
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_polygons_golden.npz")
REF_SET = "/root/reference/Dialog/dataForPlane/source_plane_registration.pcd"


def gold():
    g = np.load(GOLD)
    meta = json.loads(str(g["meta"]))
    sizes = g["sizes"]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    borders = []
    for k in range(len(sizes)):
        b = np.ones((sizes[k], 4), np.float32)
        b[:, :3] = g["vertices"][offs[k]: offs[k + 1]]
        borders.append(b)
    return g, meta, borders


def test_polygon_set_round_trip(tmp_path):
    g, meta, borders = gold()
    coeffs = np.array([g[f"coeff{k}"] for k in range(len(borders))], np.float32)
    path = str(tmp_path / "planes.pcd")
    polyset.save_polygon_set(path, borders, coeffs, 0.5)
    for suffix in ("_polySize.txt", "_polyNormal.pcd", "_polyScale.txt"):
        assert os.path.exists(path[:-4] + suffix)
    b2, n2, s2 = polyset.load_polygon_set(path)
    assert [len(b) for b in b2] == [len(b) for b in borders] == g["sizes"].tolist()
    for a, b in zip(b2, borders):
        assert a.tobytes() == b.tobytes()                   # shortest-exact printing reads back bit for bit
    assert n2.tobytes() == coeffs[:, :3].tobytes() and s2.tolist() == [0.5] * len(borders)
    head = open(path).read().splitlines()[:11]
    assert head[1] == "VERSION 0.7" and head[2] == "FIELDS x y z" and head[6] == "WIDTH 1960" and head[10] == "DATA ascii"


@pytest.mark.skipif(not os.path.exists(REF_SET), reason="needs /root/reference")
def test_reads_the_reference_polygon_files():
    borders, normals, scales = polyset.load_polygon_set(REF_SET)
    g, _, want = gold()
    assert normals is None and scales is None               # the registration variant: X.pcd + X.txt only
    assert [len(b) for b in borders] == [337, 194, 242, 182, 229, 154, 124, 212, 120, 89, 77]
    assert all(a.tobytes() == b.tobytes() for a, b in zip(borders, want))
    for b in borders:                                        # they are planar polygons: tiny spread along the normal
        c = polyset.plane_through_border(b)
        assert np.abs(b[:, :3].astype(np.float64) @ c[:3] + c[3]).max() < 0.02 * np.ptp(b[:, :3], axis=0).max()


def test_oracle_on_reference_polygons_matches_reference_source_vectors():
    g, meta, borders = gold()
    total = 0
    for k, b in enumerate(borders):
        got = O.points_in_poly(g[f"pts{k}"], g[f"coeff{k}"], b, meta["thresholds"][k], meta["seeds"][k])
        assert np.array_equal(got, g[f"inside{k}"]), k
        total += int(got.sum())
    assert total > 800
    if O.ref_lib() is not None:                              # and live against the reference's compiled source
        for k, b in enumerate(borders):
            live = O.ref_points_in_poly(g[f"pts{k}"], g[f"coeff{k}"], b, meta["thresholds"][k], meta["seeds"][k] + 1)
            assert np.array_equal(live, O.points_in_poly(g[f"pts{k}"], g[f"coeff{k}"], b, meta["thresholds"][k], meta["seeds"][k] + 1))


@pytest.mark.gpu
def test_gpu_reabsorb_on_reference_polygons():
    """All 11 reference polygons at once over the union of the golden point sets: the device claims exactly the points
    the reference's own isPointInPoly source claims (per polygon: golden vectors; across polygons: the oracle)."""
    import dialog_b200 as D
    g, meta, borders = gold()
    coeffs = np.array([g[f"coeff{k}"] for k in range(len(borders))], np.float32)
    with D.PlaneRansac(0) as pr:
        for k, b in enumerate(borders):                      # one polygon, its own threshold and seed: golden vectors
            pr.set_cloud(g[f"pts{k}"])
            cur, _, left = pr.reabsorb(coeffs[k: k + 1], [b], meta["thresholds"][k], meta["seeds"][k])
            assert np.array_equal(cur[0], np.nonzero(g[f"inside{k}"])[0]), k
            assert left == int((~g[f"inside{k}"]).sum())
        cloud = np.ascontiguousarray(np.concatenate([g[f"pts{k}"] for k in range(len(borders))]))
        t, seed = float(np.median(meta["thresholds"])), 4242
        want = O.reabsorb(cloud, coeffs, borders, t, seed)
        pr.set_cloud(cloud)
        cur, orig, left = pr.reabsorb(coeffs, borders, t, seed)
        for k in range(len(borders)):
            assert np.array_equal(cur[k], want.absorbed[k]), k
        assert left == len(want.remaining_idx) and sum(len(a) for a in cur) > 800
        assert pr.remaining().tobytes() == cloud[want.remaining_idx].tobytes()
