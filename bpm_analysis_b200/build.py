"""Builds libbpm_b200.so (hand-written CUDA for sm_100a, nvcc) and libbpm_host.so (the
compiled sequential classifier, g++) in-tree.

    python -m bpm_analysis_b200.build [--force]

The shared libraries are git-ignored but travel with the source tree; the Python
package loads them with ctypes and refuses to work without them.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libbpm_b200.so")
SOURCES = ["bpm_b200.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _newest_source_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")) and f != "bpm_host.h":     # the host library's header
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "csrc", "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed with exit code {res.returncode}")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


HOST_LIB_PATH = os.path.join(HERE, "libbpm_host.so")
HOST_SOURCES = ["classifier.cpp", "corrections.cpp"]
# -ffp-contract=off: the classifier's decisions must round exactly like CPython's float arithmetic
GXX_FLAGS = ["-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-Wall",
             "-Wextra"]


def build_host(force: bool = False) -> str:
    """libbpm_host.so: include/bpm_host.h (host-only code, no CUDA dependency)."""
    srcs = [os.path.join(CSRC, s) for s in HOST_SOURCES]
    newest = max([os.path.getmtime(f) for f in srcs] +
                 [os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "bpm_host.h"))])
    if not force and os.path.exists(HOST_LIB_PATH) and os.path.getmtime(HOST_LIB_PATH) >= newest:
        return HOST_LIB_PATH
    cmd = [os.environ.get("CXX", "g++"), *GXX_FLAGS, "-o", HOST_LIB_PATH, *srcs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"g++ failed with exit code {res.returncode}")
    if res.stderr:
        sys.stderr.write(res.stderr)
    return HOST_LIB_PATH


if __name__ == "__main__":
    print(build_host(force="--force" in sys.argv))
    print(build_native(force="--force" in sys.argv, verbose=True))
