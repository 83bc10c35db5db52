"""In-tree build of the CUDA engine (sm_100a only) and of the test oracles."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgblastn_b200.so")
SOURCES = ["scan_kernel.cu", "radix_sort.cu", "extend_kernel.cu", "gapped_kernel.cu", "traceback_kernel.cu", "lookup_build.cu", "group_sort.cu", "triage_kernel.cu", "dust_kernel.cu", "engine.cu", "hostpost.cpp", "setup.cpp", "dbfile.cpp", "dust.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_engine(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "gblastn_b200.h"))
    if not force and not _newer(LIB, deps):
        return LIB
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


def build_oracles():
    """oracle/liboracle.so always; oracle/_ref/libblastref.so only where /root/reference exists."""
    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["make", "-s", "port"], check=True, cwd=odir)
    if os.path.isdir("/root/reference/c++/src/algo/blast/core"):
        subprocess.run(["make", "-s", "-j8", "ref", "dust"], check=True, cwd=odir)
        if os.path.exists(LIB):
            subprocess.run(["make", "-s", "-j8", "shim"], check=True, cwd=odir)


if __name__ == "__main__":
    build_engine(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracles()
    print(LIB)
