"""In-tree build of the native pieces (nvcc for sm_100a; no JIT cache, the .so travels with the repo).

  libdpm_b200.so   CUDA kernels + the C ABI declared in include/dpm_b200.h
  clDPM*.so        pybind11 module with the reference's Python surface (host/ C++ classes)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libdpm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

CUDA_SOURCES = ["dpm_capi.cu", "neighbor.cu", "dpm3d.cu", "dpm2d.cu", "dpm_halo.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-ccbin", CXX,
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _deps(dirname: str) -> list[str]:
    out = []
    for root, _, files in os.walk(dirname):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp"))]
    out.append(os.path.join(HERE, "..", "include", "dpm_b200.h"))
    return out


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and _newer(LIB, _deps(CSRC)):
        return LIB
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs + ["-ldl"]
    print("[build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


def cldpm_path() -> str:
    return os.path.join(HERE, "clDPM" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_cldpm(force: bool = False) -> str:
    import pybind11

    out = cldpm_path()
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp"))
    if not force and _newer(out, _deps(HOST) + [LIB]):
        return out
    cmd = [CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-pthread", "-o", out] + srcs + [
        "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], "-I", HOST,
        "-I", os.path.join(HERE, "..", "include"), "-L", HERE, "-ldpm_b200", "-Wl,-rpath,$ORIGIN",
    ]
    print("[build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    if os.path.isdir(HOST):
        build_cldpm(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
