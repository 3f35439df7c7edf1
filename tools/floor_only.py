#!/usr/bin/env python
"""Runs stage A a few times on the C2 recording (or, with `holter` as the 4th argument, on a C4-style
4 kHz Holter recording of the given duration): a short command for ncu.

    python tools/floor_only.py [duration_sec] [repeats] [parity|fullrate] [holter]
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    from bpm_analysis_b200 import synth
    from bpm_analysis_b200.params import default_params
    from bpm_analysis_b200.runtime import StageARunner
    dur = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    p = default_params()
    p["save_filtered_wav"] = False
    if len(sys.argv) > 3:
        p["filter_mode"] = sys.argv[3]
    if len(sys.argv) > 4 and sys.argv[4] == "holter":
        pcm, sr, _ = synth.config_c4(seed=4, duration_sec=dur)
    else:
        pcm, sr, _ = synth.config_c2(seed=2, duration_sec=dur)
    A = StageARunner([len(pcm)], sr, p)
    A.upload([pcm])
    for _ in range(reps):
        A.launch()
    torch.cuda.synchronize()
    print("ok", int(A.out["peak_count"][0]))


if __name__ == "__main__":
    main()
