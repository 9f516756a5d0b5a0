"""Builds soft-body-simulation-cuda_b200/libpd_b200.so in-tree with nvcc for sm_100a.

Usage: python soft-body-simulation-cuda_b200/build.py [--force] [--verbose]
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("PD_OUT", os.path.join(HERE, "libpd_b200.so"))      # PD_OUT / PD_DEFS: experiment variants only
DEFS = os.environ.get("PD_DEFS", "").split()
RUNNER = os.path.join(HERE, "pd_run")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = os.environ.get("PD_HOSTCXX", "/usr/bin/g++")

SOURCES = ["scene.cpp", "layout.cpp", "collision.cpp", "pd_engine.cu", "pd_linear.cu", "c_api.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma,-Wall"]


def _newest_src():
    t = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    t = max(t, os.path.getmtime(os.path.join(HERE, "..", "include", "pd_b200.h")))
    return t


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_src() and (DEFS or os.path.exists(RUNNER)):
        return OUT
    objdir = os.path.join(HERE, "build" if not DEFS else "build_" + "_".join(d.strip("-D").replace("=", "") for d in DEFS))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.rsplit(".", 1)[0] + ".o")
        cmd = [NVCC] + ARCH + COMMON + DEFS + ["-x", "cu" if s.endswith(".cu") else "c++", "-c", os.path.join(CSRC, s), "-o", o]
        if s.endswith(".cu"):
            cmd += ["-Xptxas", "-v"] if verbose else []
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(o)
    cmd = [NVCC] + ARCH + ["-ccbin", HOSTCXX, "-shared", "-o", OUT] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    if not DEFS:
        # the headless C++ host driver over the C ABI (csrc/pd_run.cpp): plain g++, no CUDA headers, finds the library next to itself
        cmd = [HOSTCXX, "-std=c++17", "-O2", "-Wall", os.path.join(CSRC, "pd_run.cpp"), "-o", RUNNER, "-L" + os.path.dirname(OUT),
               "-l:" + os.path.basename(OUT), "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
