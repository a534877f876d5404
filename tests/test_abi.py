"""CPU checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports every
symbol include/plane_ransac.h declares; the product never reaches into oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "plane_ransac.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(plane_ransac_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported(lib_built):
    names = _declared_symbols()
    assert len(names) >= 25
    L = ctypes.CDLL(lib_built)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    from dialog_b200 import _lib
    assert sorted(_lib.SYMBOLS) == names


def test_abi_version_and_defaults(lib_built):
    from dialog_b200 import _lib
    L = _lib.load()
    assert L.plane_ransac_abi_version() == 3
    p = _lib.PrParams()
    L.plane_ransac_default_params(ctypes.byref(p))
    # Dialog/config.txt:29 T_dist_point_plane, :20 T_num_of_single_plane; PCL SACSegmentation defaults
    assert (p.distance_threshold, p.max_iterations, p.min_plane_size, p.probability) == (0.1, 50, 500, 0.99)
    assert (p.optimize_coefficients, p.seed, p.dot_order, p.scorer) == (1, 12345, _lib.DOT_FMA, _lib.SCORER_BRUTE)


def test_library_does_not_link_the_oracle_or_need_nccl_at_load(lib_built):
    out = subprocess.run(["ldd", lib_built], capture_output=True, text=True).stdout
    assert "oracle" not in out and "nccl" not in out
    syms = subprocess.run(["nm", "-D", lib_built], capture_output=True, text=True).stdout
    assert "orc_" not in syms


def test_product_sources_never_reference_the_oracle():
    bad = []
    for base in ("dialog_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"\boracle\b|pr_oracle|orc_", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_create_without_gpu_fails_loudly(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dialog_b200 import PlaneRansac, PlaneRansacError
    with pytest.raises(PlaneRansacError) as e:
        PlaneRansac(0)
    assert "no CPU fallback" in str(e.value)


def test_header_is_plain_c99_and_shim_compiles(tmp_path):
    """include/plane_ransac.h is the FFI surface (cgo / ctypes / JNI bind to it): it must compile as C99 on its own; the
    C++ shim (PlaneDetect-style surface incl. postProcess) must compile against it."""
    import shutil
    import subprocess
    cc = shutil.which("/usr/bin/gcc") or shutil.which("gcc")
    cxx = shutil.which("/usr/bin/g++") or shutil.which("g++")
    c_src = tmp_path / "abi.c"
    c_src.write_text('#include "plane_ransac.h"\nint main(void){ pr_params p; plane_ransac_default_params(&p); '
                     'return (int)sizeof(pr_normal) - 16 + (int)sizeof(pr_point) - 16; }\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{inc}", "-c", str(c_src), "-o", str(tmp_path / "abi.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp_src = tmp_path / "shim.cpp"
    cpp_src.write_text('#include "PlaneDetectRansac.h"\n'
                       'bool use(plane_detect_ransac::PlaneDetectRansac& d, std::vector<plane_detect_ransac::PointXYZ>& c) {\n'
                       '  std::vector<plane_detect_ransac::PlaneRecord> p;\n'
                       '  std::vector<std::vector<plane_detect_ransac::PointXYZ>> b;\n'
                       '  return d.detect(c, p) && d.postProcess(c, p, b, 1u);\n}\n')
    r = subprocess.run([cxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", f"-I{inc}", "-c", str(cpp_src), "-o", str(tmp_path / "shim.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_pcl_overload_compiles_and_links_against_stub_pcl(lib_built, tmp_path):
    """The PLANE_RANSAC_WITH_PCL overload (pcl::PointCloud<pcl::PointXYZ>::Ptr in, pcl::ModelCoefficients +
    pcl::PointIndices out — the surface north_star names) over a 30-line stand-in for the four PCL headers it includes:
    PCL itself is not in this image.  tests/test_gpu_parity.py runs the resulting program on the GPU box."""
    import shutil
    import subprocess
    cxx = shutil.which("/usr/bin/g++") or shutil.which("g++")
    out = tmp_path / "pcl_overload_check"
    r = subprocess.run([cxx, "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", f"-I{os.path.join(ROOT, 'include')}",
                        f"-I{os.path.join(ROOT, 'tests', 'pcl_stub')}", os.path.join(ROOT, "tests", "pcl_overload_check.cpp"), "-o", str(out),
                        f"-L{os.path.join(ROOT, 'dialog_b200')}", "-lplane_ransac", f"-Wl,-rpath,{os.path.join(ROOT, 'dialog_b200')}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out.exists()
