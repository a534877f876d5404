"""The reference's on-disk polygon set: what PCLViewer::on_savePolyDataAction_triggered writes
(Dialog/PCLViewer.cpp:1341-1396) and the loader at Dialog/PCLViewer.cpp:1005-1100 reads back:

    X.pcd              ASCII PCD of every border vertex, polygon after polygon (pcl::io::savePCDFileASCII)
    X_polySize.txt     one line per polygon: its vertex count
    X_polyNormal.pcd   ASCII PCD of pcl::Normal, one per polygon: the plane normal (Plane::coeff.values[0..3))
    X_polyScale.txt    one line per polygon: r_for_estimate_normal at save time

and the two-file variant the registration path keeps in Dialog/dataForPlane (X.pcd + X.txt with the vertex counts).
Host-side I/O only: the polygons are the input of PlaneRansac.reabsorb and the hand-off to the polygon half of the
reference.  Floats are printed with the shortest representation that reads back to the same float32."""
from __future__ import annotations

import os

import numpy as np

from .pcd import read_pcd_xyz


def _fmt(v: float) -> str:
    f = np.float32(v)
    for prec in range(1, 10):  # PCL prints with up to 8-9 significant digits; take the shortest exact one
        s = np.format_float_positional(f, precision=prec, unique=False, fractional=False, trim="-")
        if np.float32(s) == f:
            return s
    return repr(float(f))


def _write_ascii_pcd(path: str, fields, rows) -> None:
    rows = np.asarray(rows, np.float32).reshape(-1, len(fields))
    n = rows.shape[0]
    with open(path, "w") as f:
        f.write("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n")
        f.write("FIELDS " + " ".join(fields) + "\n")
        f.write("SIZE " + " ".join(["4"] * len(fields)) + "\n")
        f.write("TYPE " + " ".join(["F"] * len(fields)) + "\n")
        f.write("COUNT " + " ".join(["1"] * len(fields)) + "\n")
        f.write(f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA ascii\n")
        for r in rows:
            f.write(" ".join(_fmt(v) for v in r) + "\n")


def _read_ascii_fields(path: str, wanted) -> np.ndarray:
    hdr, rows = {}, []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            if "DATA" not in hdr:
                key, *vals = line.split()
                hdr[key.upper()] = vals
                continue
            rows.append(line.split())
    if hdr.get("DATA", [""])[0].lower() != "ascii":
        raise ValueError(f"{path}: only ASCII polygon-normal files are written by the reference")
    cols = [hdr["FIELDS"].index(w) for w in wanted]
    return np.array([[np.float32(r[c]) for c in cols] for r in rows], np.float32).reshape(-1, len(wanted))


def save_polygon_set(path_pcd: str, borders, coeffs, scale: float) -> None:
    """borders: list of (nb, 3|4) vertex arrays (Plane::border); coeffs: (P, >=3) plane coefficients; scale: the
    r_for_estimate_normal the reference records per polygon."""
    if not path_pcd.endswith(".pcd"):
        raise ValueError("the reference derives the side files by dropping a 4-character extension: use X.pcd")
    stem = path_pcd[:-4]
    verts = np.concatenate([np.asarray(b, np.float32)[:, :3] for b in borders]) if len(borders) else np.zeros((0, 3), np.float32)
    _write_ascii_pcd(path_pcd, ("x", "y", "z"), verts)
    with open(stem + "_polySize.txt", "w") as f:
        for b in borders:
            f.write(f"{len(b)}\n")
    co = np.asarray(coeffs, np.float32).reshape(len(borders), -1)
    # pcl::Normal prints normal_x normal_y normal_z curvature; the reference leaves curvature unset (written as 0 here)
    _write_ascii_pcd(stem + "_polyNormal.pcd", ("normal_x", "normal_y", "normal_z", "curvature"),
                     np.concatenate([co[:, :3], np.zeros((len(borders), 1), np.float32)], 1))
    with open(stem + "_polyScale.txt", "w") as f:
        for _ in borders:
            f.write(f"{_fmt(scale)}\n")


def load_polygon_set(path_pcd: str):
    """(borders, normals, scales) from X.pcd + X_polySize.txt [+ X_polyNormal.pcd, X_polyScale.txt], or from the
    registration variant X.pcd + X.txt (normals and scales are None then).  borders: list of (nb, 4) float32 rows
    (x, y, z, 1) like pcl::PointXYZ."""
    stem = path_pcd[:-4]
    verts = read_pcd_xyz(path_pcd)
    size_file = stem + "_polySize.txt" if os.path.exists(stem + "_polySize.txt") else stem + ".txt"
    with open(size_file) as f:
        sizes = [int(line) for line in f if line.strip()]  # atoi per line, as the reference's loader
    if sum(sizes) > len(verts):
        raise ValueError(f"{size_file}: {sum(sizes)} vertices listed, {len(verts)} in {path_pcd}")
    borders, at = [], 0
    for s in sizes:
        borders.append(np.ascontiguousarray(verts[at: at + s]))
        at += s
    normals = scales = None
    if os.path.exists(stem + "_polyNormal.pcd"):
        normals = _read_ascii_fields(stem + "_polyNormal.pcd", ("normal_x", "normal_y", "normal_z"))
    if os.path.exists(stem + "_polyScale.txt"):
        with open(stem + "_polyScale.txt") as f:
            scales = np.array([np.float32(line) for line in f if line.strip()], np.float32)
    return borders, normals, scales


def plane_through_border(border) -> np.ndarray:
    """Least-squares plane (a, b, c, d), unit normal, through a polygon's vertices — what the reference recomputes from
    points_set with pcl::computePointNormal when it reloads a polygon set (Dialog/PCLViewer.cpp:1076-1100)."""
    v = np.asarray(border, np.float64)[:, :3]
    c = v.mean(0)
    w, e = np.linalg.eigh(np.cov((v - c).T))
    n = e[:, 0]
    return np.array([n[0], n[1], n[2], -float(n @ c)], np.float32)
