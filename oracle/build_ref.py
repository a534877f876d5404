"""Compiles the reference's OWN source of the re-absorption predicate as a checker: oracle/_ref/libdialog_ref.so.

TEST INFRASTRUCTURE ONLY.  The function bodies of distP2P, projPoint2Plane, getInfoBetPointAndPlane,
isBothLineSegsIntersect and isPointInPoly are read at build time from /root/reference/Dialog/PlaneDetect.h where
they lie (GBK -> UTF-8, located by signature + brace matching, no line numbers), written to a temporary file
outside the repo, compiled over oracle/ref_shim.h (stand-ins for the PCL/Eigen types; PCL and Eigen themselves are
absent from this image) and the temporary file is deleted.  Only the shared library lands in oracle/_ref/
(git-ignored; it travels to the GPU box).  No reference source is copied into the repository.

    python oracle/build_ref.py            # no-op (exit 0) when /root/reference is absent
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Dialog/PlaneDetect.h"
OUT = os.path.join(HERE, "_ref", "libdialog_ref.so")

SIGNATURES = [
    r"inline\s+float\s+distP2P\s*\(",
    r"void\s+projPoint2Plane\s*\(",
    r"bool\s+isPointInPoly\s*\(",
    r"bool\s+isBothLineSegsIntersect\s*\(",
    r"void\s+getInfoBetPointAndPlane\s*\(",
]

WRAPPERS = r"""
extern "C" {
int ref_segs_intersect(const float* pa, const float* pb, const float* pc, const float* pd) {
  pcl::PointXYZ a{pa[0], pa[1], pa[2], 1.f}, b{pb[0], pb[1], pb[2], 1.f}, c{pc[0], pc[1], pc[2], 1.f}, d{pd[0], pd[1], pd[2], 1.f}, x;
  return isBothLineSegsIntersect(a, b, c, d, x) ? 1 : 0;
}
// points / border: 4 floats per point (pcl::PointXYZ layout); out[i] = isPointInPoly(points[i], plane)
void ref_points_in_poly(const float* points, size_t n, const float* coeff, const float* border, int nb, float t,
                        unsigned seed, unsigned char* out) {
  PointCloudT cloud;
  cloud.points.resize(nb);
  for (int j = 0; j < nb; ++j) cloud.points[j] = pcl::PointXYZ{border[4 * j], border[4 * j + 1], border[4 * j + 2], 1.f};
  Plane plane;
  plane.border = &cloud;
  plane.points_set = nullptr;
  plane.coeff.values.assign(coeff, coeff + 4);
  T_dist_point_plane = t;
  ref_time_value = seed;
  for (size_t i = 0; i < n; ++i) {
    pcl::PointXYZ p{points[4 * i], points[4 * i + 1], points[4 * i + 2], 1.f};
    out[i] = isPointInPoly(p, plane) ? 1 : 0;
  }
}
void ref_project(const float* p, const float* coeff, float* out3, float* dist) {
  pcl::PointXYZ s{p[0], p[1], p[2], 1.f}, d;
  Eigen::Vector4f c;
  c << coeff[0], coeff[1], coeff[2], coeff[3];
  getInfoBetPointAndPlane(s, c, *dist, d);
  out3[0] = d.x; out3[1] = d.y; out3[2] = d.z;
}
}
"""


def extract(text: str, sig: str) -> str:
    """The definition (signature ... matching closing brace) of the function whose signature matches sig."""
    for m in re.finditer(sig, text):
        start = text.rfind("\n", 0, m.start()) + 1
        brace = text.find("{", m.end())
        semi = text.find(";", m.end())
        if brace < 0 or (0 <= semi < brace):
            continue  # a prototype
        depth, i = 0, brace
        while True:
            ch = text[i]
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
                if depth == 0:
                    return text[start:i + 1]
            i += 1
    raise RuntimeError(f"definition not found: {sig}")


def build(force: bool = False) -> str | None:
    if not os.path.exists(REF):
        return OUT if os.path.exists(OUT) else None
    srcs = [REF, os.path.join(HERE, "ref_shim.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    text = open(REF, encoding="gbk", errors="replace").read()
    parts = ['#include "ref_shim.h"\n'] + [extract(text, s) + "\n" for s in SIGNATURES] + [WRAPPERS]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cxx = shutil.which("/usr/bin/g++") or shutil.which("g++")
    with tempfile.TemporaryDirectory() as tmp:
        tu = os.path.join(tmp, "dialog_ref_tu.cpp")
        with open(tu, "w", encoding="utf-8") as f:
            f.write("".join(parts))
        cmd = [cxx, "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-w", f"-I{HERE}", tu, "-o", OUT]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building oracle/_ref/libdialog_ref.so failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
