"""PCD reader on the boundary: what pcl::io::loadPCDFile gives the reference's callers
(Dialog/PCLViewer.cpp:80-89) for an XYZ cloud.  The reader itself is plane_ransac_load_pcd in the C ABI
(csrc/pr_pcd.cpp: PCD v0.7 ascii / binary / binary_compressed, x y z anywhere in the record, other fields skipped); this
module only hands its page-locked result to numpy as rows laid out like pcl::PointXYZ (x, y, z, 1)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def read_pcd_xyz(path: str) -> np.ndarray:
    L = _lib.load()
    pts, n = C.c_void_p(None), C.c_size_t(0)
    rc = L.plane_ransac_load_pcd(str(path).encode(), C.byref(pts), C.byref(n))
    if rc != 0:
        raise ValueError(L.plane_ransac_last_error().decode("utf-8", "replace"))
    try:
        if n.value == 0:
            return np.ones((0, 4), np.float32)
        buf = (C.c_float * (4 * n.value)).from_address(pts.value)
        return np.frombuffer(buf, np.float32).reshape(n.value, 4).copy()
    finally:
        L.plane_ransac_host_free(pts)
