#!/usr/bin/env python
"""Per-phase cycle counts of k_rolling_floor_blk (diagnostic build with -DBPM_DEBUG_COUNTERS).

    python tools/floor_phases.py            # on a GPU box; builds libbpm_b200_dbg.so if missing

Phases (RB_TICK indices): 0 stage samples, 1 pivot probes, 2 splitter sample + bitonic sort,
3 bucket ids, 4 scan + scatter, 5 in-bucket order + inverse permutation, 6 coarse table,
7 slide.  Counter 15 = CTAs that had to repeat with pivot = +inf.
"""
import ctypes as C
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
DBG = os.path.join(REPO, "bpm_analysis_b200", "libbpm_b200_dbg.so")


def build():
    from bpm_analysis_b200.build import CSRC, NVCC_FLAGS, SOURCES, find_nvcc
    cmd = [find_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], "-shared", "-DBPM_DEBUG_COUNTERS", "-o", DBG,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    subprocess.run(cmd, check=True)


def main():
    if "--build-only" in sys.argv or not os.path.exists(DBG):
        build()
        if "--build-only" in sys.argv:
            return
    os.environ["BPM_B200_LIB"] = DBG
    import numpy as np
    import torch
    from bpm_analysis_b200 import _native, synth
    from bpm_analysis_b200.params import default_params
    from bpm_analysis_b200.runtime import StageARunner
    lib = _native.load_library(DBG)
    lib.bpm_debug_counters.restype = C.c_int
    lib.bpm_debug_counters.argtypes = [C.c_void_p, C.c_int]
    p = default_params()
    p["save_filtered_wav"] = False
    dur = float(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1][0] != "-" else 3600.0
    if len(sys.argv) > 2 and sys.argv[2] in ("parity", "fullrate"):
        p["filter_mode"] = sys.argv[2]
    pcm, sr, _ = synth.config_c2(seed=2, duration_sec=dur)
    A = StageARunner([len(pcm)], sr, p)
    A.upload([pcm])
    A.launch()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    lib.bpm_debug_counters(buf, 1)
    A.launch()
    torch.cuda.synchronize()
    lib.bpm_debug_counters(buf, 1)
    v = np.array(list(buf), dtype=np.float64)
    if hasattr(lib, "bpm_debug_counters_scan"):
        lib.bpm_debug_counters_scan.restype = C.c_int
        lib.bpm_debug_counters_scan.argtypes = [C.c_void_p, C.c_int]
        sb = (C.c_ulonglong * 16)()
        lib.bpm_debug_counters_scan(sb, 1)
        A.launch()
        torch.cuda.synchronize()
        lib.bpm_debug_counters_scan(sb, 1)
        sv = np.array(list(sb), dtype=np.float64)
        sn = ["constants", "fill", "sweep1+warp scan", "publish", "look-back", "prefixes", "sweep2"]
        for dname, o in (("fwd", 0), ("bwd", 8)):
            tot = sv[o:o + 7].sum()
            print(f"k_scan {dname}: thread-0 cycles summed over CTAs {tot:.3e}")
            for i, nme in enumerate(sn):
                print(f"  {nme:18s} {sv[o + i]:.3e}  {100 * sv[o + i] / max(tot, 1):5.1f} %")
    names = ["stage", "probes", "splitters", "bucket ids", "scan+scatter", "in-bucket+perm", "table", "slide"]
    tot = v[:8].sum()
    print(f"cycles summed over CTAs of 2 launches: {tot:.3e}; repeated CTAs: {int(v[15])}")
    for i, nme in enumerate(names):
        print(f"  {nme:16s} {v[i]:.3e}  {100 * v[i] / tot:5.1f} %")


if __name__ == "__main__":
    main()
