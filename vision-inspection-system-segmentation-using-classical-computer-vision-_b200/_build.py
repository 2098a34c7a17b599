"""Build the C-ABI shared library in-tree with nvcc (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libvi_b200.so")
CHECKED_LIB_PATH = os.path.join(PKG_DIR, "libvi_b200_checked.so")     # same sources with -DVI_CHECKED=1 (bounds self-checks)
SOURCES = [os.path.join(PKG_DIR, "csrc", "vi_api.cu")]
HEADERS = [os.path.join(PKG_DIR, "csrc", n) for n in ("vi_device.cuh", "vi_ccl.cuh", "vi_pipeline.cuh", "vi_rank.cuh", "vi_canny.cuh", "vi_unit.cuh", "vi_ingest.cuh")]
HEADERS.append(os.path.join(ROOT, "include", "vi_b200.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def needs_build(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False, checked=False):
    out = CHECKED_LIB_PATH if checked else LIB_PATH
    if not force and not needs_build(out):
        return out
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DVI_CHECKED=1"] if checked else []) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out
