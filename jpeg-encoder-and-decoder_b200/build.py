"""In-tree build of the B200 encode path.

  libjpgb200.so    CUDA kernels + C ABI (include/jpgb200.h), sm_100a only
  libjpgenc.so     host-side C mirror of the reference's encoder API (host/*.c)
  jpg              the reference-style driver executable (host/jpg_driver.c)

nvcc cross-compiles without a GPU.  Built artefacts stay in-tree (git-ignored) so they
travel to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libjpgb200.so")
HOSTLIB = os.path.join(HERE, "libjpgenc.so")
DRIVER = os.path.join(HERE, "jpg")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "--extended-lambda",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return out


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libjpgb200.so cannot be built (there is no CPU fallback)")
    return p


def build_cuda(force=False, verbose=False):
    srcs = _sources(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "jpgb200.h")]
    if force or _newer(LIB, srcs):
        extra = os.environ.get("JPGB200_NVCC_EXTRA", "").split()  # e.g. -DK1C_NC=2 for A/B builds of a kernel variant
        if os.environ.get("JPGB200_AB", "0") == "1":
            extra.append("-DJPGB200_AB")  # the older K1 forms and K35: cross-checks, not product (include/jpgb200.h)
        cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "jpgb200.cu")]
        subprocess.check_call(cmd)
    return LIB


def build_host(force=False):
    srcs = _sources(HOST, (".c", ".h")) + [os.path.join(ROOT, "include", "jpgb200.h")]
    if not os.path.isdir(HOST):
        return None
    cfiles = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".c") and f != "jpg_driver.c"]
    if force or _newer(HOSTLIB, srcs + [LIB]):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-fPIC", "-shared", "-I", HOST, "-I", os.path.join(ROOT, "include"),
                               "-o", HOSTLIB] + cfiles + ["-L", HERE, "-ljpgb200", "-Wl,-rpath,$ORIGIN", "-lm"])
    drv = os.path.join(HOST, "jpg_driver.c")
    if os.path.exists(drv) and (force or _newer(DRIVER, srcs + [HOSTLIB])):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-I", HOST, "-I", os.path.join(ROOT, "include"), "-o", DRIVER, drv,
                               "-L", HERE, "-ljpgenc", "-ljpgb200", "-Wl,-rpath,$ORIGIN", "-lm"])
    return HOSTLIB


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
