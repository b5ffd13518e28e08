"""Build recipe for the native library: plain nvcc, sm_100a only, in-tree output.

    python -m yolo_dual_b200.build [--force] [--verbose]

The reference builds a torch CUDAExtension with no arch flags
(models/ops_dcnv3/setup.py:39-59, ~50 s); this library has no torch headers and
compiles in ~75 s.  The .so is git-ignored but travels to the GPU box in-tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdcnv3_b200.so")
SOURCES = ["dcnv3_b200.cu"]
HEADERS = ["dcnv3_common.cuh", "dcnv3_kernels.cuh", "dcnv3_bwd_tile.cuh", "dcnv3_imat.cuh", "dcnv3_win.cuh", os.path.join(ROOT, "include", "dcnv3_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # no --split-compile: it halves the build time but ptxas then compiles the imat backward's
    # non-inlined point functions apart from their kernel (standard ABI instead of a tailored calling
    # convention) and the kernel runs ~10 % slower (measured: P3 backward 270 us vs 243 us)
    "-shared", "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
]


# second, small library: the fused CE + Dice loss of the seg trainers (include/segloss_b200.h)
SEGLOSS_LIB = os.path.join(CSRC, "libsegloss_b200.so")
SEGLOSS_DEPS = [os.path.join(CSRC, "segloss_b200.cu"), os.path.join(ROOT, "include", "segloss_b200.h")]


# third: fused training-mode BatchNorm2d + SiLU of the Conv block (include/bnact_b200.h)
BNACT_LIB = os.path.join(CSRC, "libbnact_b200.so")
BNACT_DEPS = [os.path.join(CSRC, "bnact_b200.cu"), os.path.join(ROOT, "include", "bnact_b200.h")]


# fourth: NHWC nearest / bilinear resize of the seg heads (include/resize_b200.h)
RESIZE_LIB = os.path.join(CSRC, "libresize_b200.so")
RESIZE_DEPS = [os.path.join(CSRC, "resize_b200.cu"), os.path.join(ROOT, "include", "resize_b200.h")]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the dcnv3_b200 library cannot be built here")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + \
           [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def _run_nvcc(out: str, sources, verbose: bool):
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + list(sources)
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a gcc wrapper nvcc should not pick up
    env.pop("CXX", None)
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout)
    if r.returncode:
        raise RuntimeError("nvcc failed:\n" + r.stdout[-4000:])


def build_segloss(force: bool = False, verbose: bool = False) -> str:
    if force or not os.path.exists(SEGLOSS_LIB) or \
            any(os.path.getmtime(d) > os.path.getmtime(SEGLOSS_LIB) for d in SEGLOSS_DEPS):
        _run_nvcc(SEGLOSS_LIB, SEGLOSS_DEPS[:1], verbose)
    return SEGLOSS_LIB


def build_bnact(force: bool = False, verbose: bool = False) -> str:
    if force or not os.path.exists(BNACT_LIB) or \
            any(os.path.getmtime(d) > os.path.getmtime(BNACT_LIB) for d in BNACT_DEPS):
        _run_nvcc(BNACT_LIB, BNACT_DEPS[:1], verbose)
    return BNACT_LIB


def build_resize(force: bool = False, verbose: bool = False) -> str:
    if force or not os.path.exists(RESIZE_LIB) or \
            any(os.path.getmtime(d) > os.path.getmtime(RESIZE_LIB) for d in RESIZE_DEPS):
        _run_nvcc(RESIZE_LIB, RESIZE_DEPS[:1], verbose)
    return RESIZE_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    build_segloss(force, verbose)
    build_bnact(force, verbose)
    build_resize(force, verbose)
    if not force and not _stale():
        return LIB
    _run_nvcc(LIB, [os.path.join(CSRC, s) for s in SOURCES], verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
