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
SOURCES = ["filter.cu", "sosfilt.cu", "contract.cu", "select.cu", "peaks.cu", "floor.cu", "metrics.cu", "shard.cu", "pipeline.cu"]
OBJ_DIR = os.path.join(CSRC, "_obj")            # git-ignored (*.o)
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
# ptxas -v of every kernel (registers, shared memory, spills): tracked, so that the binary the GPU box
# ran can be tied to a register budget (profiles/ cites it)
BUILD_LOG = os.path.join(CSRC, "build.log")


def _header_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")) and f != "bpm_host.h":     # the host library's header
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def build_native(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """One object per .cu (compiled in parallel; every translation unit launches only its own kernels,
    so no relocatable device code), linked into libbpm_b200.so.  Only stale objects are recompiled."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc, hdr = find_nvcc(), _header_mtime()
    flag_tag = os.path.join(OBJ_DIR, "flags.txt")
    flags = [*NVCC_FLAGS, *extra_flags]
    if not os.path.exists(flag_tag) or open(flag_tag).read() != " ".join(flags):
        force = True
    jobs = []
    for src in SOURCES:
        sp, op = os.path.join(CSRC, src), os.path.join(OBJ_DIR, src[:-3] + ".o")
        if force or not os.path.exists(op) or os.path.getmtime(op) < max(os.path.getmtime(sp), hdr):
            jobs.append((src, [nvcc, *flags, "-c", sp, "-o", op]))
    objs = [os.path.join(OBJ_DIR, s[:-3] + ".o") for s in SOURCES]
    if not jobs and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(o) for o in objs):
        return LIB_PATH
    logs = {}
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for (src, cmd), res in zip(jobs, ex.map(lambda j: subprocess.run(j[1], capture_output=True, text=True), jobs)):
                logs[src] = " ".join(cmd) + "\n" + res.stdout + res.stderr
                if res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                    raise RuntimeError(f"nvcc failed on {src} with exit code {res.returncode}")
                if verbose:
                    sys.stderr.write(res.stderr)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"link failed with exit code {res.returncode}")
    with open(flag_tag, "w") as fh:
        fh.write(" ".join(flags))
    # the log keeps one section per source; sections of objects that were not rebuilt are carried over
    old = {}
    if os.path.exists(BUILD_LOG):
        cur = None
        for line in open(BUILD_LOG):
            if line.startswith("### "):
                cur = line[4:].strip()
                old[cur] = ""
            elif cur is not None:
                old[cur] += line
    old.update(logs)
    with open(BUILD_LOG, "w") as fh:
        for src in SOURCES:
            if src in old:
                fh.write(f"### {src}\n{old[src]}")
    return LIB_PATH


HOST_LIB_PATH = os.path.join(HERE, "libbpm_host.so")
HOST_SOURCES = ["classifier.cpp", "corrections.cpp", "ingest.cpp", "reports.cpp"]
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
