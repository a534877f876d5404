"""Builds libplane_ransac.so (CUDA kernels for sm_100a + the C ABI) in-tree with nvcc / g++.

    python -m dialog_b200.build [--force]

The library is built next to this file so that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libplane_ransac.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
# the image exports CXX=/opt/gcc/bin/g++; the system g++ is the one nvcc uses as host compiler
CXX = shutil.which("/usr/bin/g++") or shutil.which("g++")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # FP contraction only where an fma intrinsic is written (bit-exact parity)
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-Wextra",
             f"-I{CUDA_HOME}/include"]

SOURCES = [("pr_kernels.cu", "nvcc"), ("pr_chain.cu", "nvcc"), ("pr_reabsorb.cu", "nvcc"), ("pr_p2p.cu", "nvcc"), ("pr_normals.cu", "nvcc"), ("pr_host.cpp", "cxx"), ("pr_pcd.cpp", "cxx"), ("pr_api.cpp", "cxx")]
HEADERS = ["pr_kernels.h", "pr_host.hpp", "pr_math.h", "pr_draw.h", os.path.join("..", "..", "include", "plane_ransac.h")]


def _newer(src_paths, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in src_paths)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    for name, kind in SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ, os.path.splitext(name)[0] + ".o")
        objs.append(obj)
        if force or _newer([src] + hdrs, obj):
            if kind == "nvcc":
                cmd = [NVCC, *NVCC_FLAGS, "-Xptxas", "-v", "-c", src, "-o", obj]
            else:
                cmd = [CXX, *CXX_FLAGS, "-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"compiling {name} failed")
            if kind == "nvcc":
                with open(os.path.join(OBJ, "ptxas_" + os.path.splitext(name)[0] + ".log"), "w") as f:
                    f.write(r.stderr)
    if force or _newer(objs, LIB):
        cmd = [CXX, "-shared", "-o", LIB, *objs, f"-L{CUDA_HOME}/lib64", "-lcudart_static", "-ldl", "-lrt",
               "-lpthread", "-Wl,--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("linking libplane_ransac.so failed")
    build_demo(force)
    return LIB


DEMO_SRC = os.path.join(os.path.dirname(HERE), "examples", "plane_detect_demo.cpp")
DEMO_BIN = os.path.join(os.path.dirname(HERE), "examples", "plane_detect_demo")
SHIM = os.path.join(os.path.dirname(HERE), "include", "PlaneDetectRansac.h")


def build_demo(force: bool = False) -> str:
    """The C++ program that drives include/PlaneDetectRansac.h (the PlaneDetect-style shim)."""
    if force or _newer([DEMO_SRC, SHIM, LIB], DEMO_BIN):
        cmd = [CXX, "-O2", "-std=c++17", "-Wall", "-Wextra", f"-I{os.path.join(os.path.dirname(HERE), 'include')}",
               DEMO_SRC, "-o", DEMO_BIN, f"-L{HERE}", "-lplane_ransac", "-Wl,-rpath,$ORIGIN/../dialog_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("building examples/plane_detect_demo failed")
    return DEMO_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
