#!/usr/bin/env python
"""Counters of k_distance_tiles on the C2 / C4 envelopes (diagnostic build, -DBPM_DEBUG_COUNTERS):
tiles, rounds per tile, candidates left to the global finish, cycles per phase.

    python tools/peaks_probe.py [c2|c4]      # needs bpm_analysis_b200/libbpm_b200_dbg.so (tools/floor_phases.py --build-only)
"""
import ctypes as C
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
DBG = os.path.join(REPO, "bpm_analysis_b200", "libbpm_b200_dbg.so")


def main():
    os.environ["BPM_B200_LIB"] = DBG
    import numpy as np
    import torch
    from bpm_analysis_b200 import _native, synth
    from bpm_analysis_b200.params import default_params
    from bpm_analysis_b200.runtime import StageARunner
    lib = _native.load_library(DBG)
    lib.bpm_debug_counters_peaks.restype = C.c_int
    lib.bpm_debug_counters_peaks.argtypes = [C.c_void_p, C.c_int]
    p = default_params()
    p["save_filtered_wav"] = False
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    pcm, sr, _ = synth.config_c2(seed=2) if which == "c2" else synth.config_c4(seed=4, duration_sec=86400.0)
    A = StageARunner([len(pcm)], sr, p)
    A.upload([pcm])
    A.launch()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    lib.bpm_debug_counters_peaks(buf, 1)
    A.launch()
    torch.cuda.synchronize()
    lib.bpm_debug_counters_peaks(buf, 1)
    v = list(buf)
    names = ["tiles", "rounds (sum)", "rounds (max of a tile)", "pending after tiles", "global-finish rounds",
             "candidates (sum over launches)", "cycles staging", "cycles rounds", "cycles write-back"]
    print(which, "M =", A.total_m, "troughs", int(A.out["trough_count"][0]), "peaks", int(A.out["peak_count"][0]))
    for n, x in zip(names, v):
        print(f"  {n:32s} {x}")
    if v[0]:
        print(f"  rounds per tile {v[1] / v[0]:.1f}; cycles per tile: staging {v[6] / v[0]:.0f}, rounds {v[7] / v[0]:.0f}, "
              f"write-back {v[8] / v[0]:.0f}")


if __name__ == "__main__":
    main()
