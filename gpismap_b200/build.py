"""Build recipes for the in-tree native code (nvcc cross-compiles sm_100a without a GPU).

    python -m gpismap_b200.build            # libgpis_b200.so (+ host library when present)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB_CUDA = os.path.join(HERE, "libgpis_b200.so")
LIB_HOST = os.path.join(HERE, "libgpismap_host.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # The reference is built without FMA contraction (SSE2, mex/make_GPisMap3.m:15); geometry tests and
    # covariance entries must round the same way. Hot loops use explicit fmaf().
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "gpis_b200.h")]
    if not force and not _newer(LIB_CUDA, srcs):
        return LIB_CUDA
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_CUDA, os.path.join(CSRC, "gpis_b200.cu")]
    subprocess.check_call(cmd)
    return LIB_CUDA


def build_host(force=False):
    if not os.path.isdir(HOST):
        return None
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST))] + [os.path.join(ROOT, "include", "gpis_b200.h")]
    srcs += [os.path.join(ROOT, "include", "gpismap", f) for f in sorted(os.listdir(os.path.join(ROOT, "include", "gpismap")))]
    if not force and not _newer(LIB_HOST, srcs):
        return LIB_HOST
    cpps = [s for s in srcs if s.endswith(".cpp")]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
           "-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-o", LIB_HOST] + cpps + \
          ["-L" + HERE, "-lgpis_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return LIB_HOST


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB_CUDA)
