"""Minimal PCD v0.7 reader for the boundary: what pcl::io::loadPCDFile gives the
reference's callers (Dialog/PCLViewer.cpp:80-89) for an XYZ cloud — ASCII and uncompressed binary,
x/y/z float32 fields anywhere in the record, other fields skipped.  Returns rows laid out like
pcl::PointXYZ (x, y, z, 1), ready for PlaneRansac.set_cloud."""
from __future__ import annotations

import numpy as np

_NP = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4", ("U", 8): "<u8",
       ("I", 1): "i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8"}


def read_pcd_xyz(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        raw = f.read()
    hdr = {}
    pos = 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, *vals = line.split()
        hdr[key.upper()] = vals
        if key.upper() == "DATA":
            break
    fields = hdr.get("FIELDS", [])
    if not all(a in fields for a in ("x", "y", "z")):
        raise ValueError(f"{path}: no x y z fields")
    sizes = [int(v) for v in hdr["SIZE"]]
    types = hdr["TYPE"]
    counts = [int(v) for v in hdr.get("COUNT", ["1"] * len(fields))]
    n = int(hdr["POINTS"][0]) if "POINTS" in hdr else int(hdr["WIDTH"][0]) * int(hdr["HEIGHT"][0])
    kind = hdr["DATA"][0].lower()
    out = np.ones((n, 4), np.float32)
    if kind == "ascii":
        cols, c = {}, 0
        for name, cnt in zip(fields, counts):
            cols[name] = c
            c += cnt
        body = raw[pos:].split()
        table = np.array(body[: n * c], dtype=object).reshape(n, c)
        for j, a in enumerate("xyz"):
            out[:, j] = table[:, cols[a]].astype(np.float64).astype(np.float32) if types[fields.index(a)] != "F" else \
                np.array([np.float32(v) for v in table[:, cols[a]]], np.float32)
    elif kind == "binary":
        dt = []
        for name, sz, ty, cnt in zip(fields, sizes, types, counts):
            if (ty, sz) not in _NP:
                raise ValueError(f"{path}: unsupported field type {ty}{sz}")
            dt.append((name, _NP[(ty, sz)], (cnt,)) if cnt > 1 else (name, _NP[(ty, sz)]))
        rec = np.frombuffer(raw, dtype=np.dtype(dt), count=n, offset=pos)
        for j, a in enumerate("xyz"):
            out[:, j] = rec[a].astype(np.float32)
    else:
        raise ValueError(f"{path}: DATA {kind} is not supported (binary_compressed needs LZF)")
    return out
