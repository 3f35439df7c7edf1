#!/usr/bin/env python
"""Where a drop-in chain's wall time goes (C2 recording): the ten frontend.* calls of
analyze_wav_file's order, timed one by one with the host clock, averaged over `reps` steps.

    python tools/dropin_profile.py [reps]
"""
import json
import logging
import os
import sys
import time

import numpy as np

os.environ.setdefault("BPM_DROPIN_TRACE", "1")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    from bpm_analysis_b200 import frontend, synth
    from bpm_analysis_b200.dropin import dropin
    from bpm_analysis_b200.params import default_params
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    logging.getLogger().setLevel(logging.ERROR)
    p = default_params()
    p["save_filtered_wav"] = False
    pcm, sr, beats = synth.config_c2(seed=2)
    pcm = torch.from_numpy(pcm).pin_memory().numpy()
    svc = dropin()
    rate = 301
    bi = synth.beats_to_envelope_indices(beats, rate)
    prelim = bi[(np.arange(len(bi)) % 9) != 4]
    acc = {}

    class Clf:
        pass

    def tick(name, t0):
        t1 = time.perf_counter()
        acc[name] = acc.get(name, 0.0) + (t1 - t0)
        return t1

    for rep in range(reps + 3):
        if rep == 3:
            acc.clear()
            svc.trace.clear()
        t = time.perf_counter()
        svc.forget()
        t = tick("forget", t)
        env, r, _, _ = frontend.preprocess_pcm(pcm, sr, p, want_filtered=False)
        t = tick("preprocess_pcm", t)
        floor, troughs = frontend._calculate_dynamic_noise_floor(env, r, p)
        t = tick("_calculate_dynamic_noise_floor", t)
        clf = Clf()
        clf.audio_envelope, clf.sample_rate, clf.params = env, r, p
        st1 = frontend._initialize_state(clf, None, floor, troughs)
        t = tick("_initialize_state #1", t)
        ps, pt = frontend.calculate_bpm_series(prelim, r, p)
        t = tick("calculate_bpm_series (prelim)", t)
        frontend.find_recovery_phase(ps, pt, p)
        t = tick("find_recovery_phase", t)
        st2 = frontend._initialize_state(clf, 80.0, floor, troughs)
        t = tick("_initialize_state #2", t)
        sm, bt = frontend.calculate_bpm_series(bi, r, p)
        t = tick("calculate_bpm_series", t)
        frontend.find_major_hr_inclines(sm)
        t = tick("find_major_hr_inclines", t)
        frontend.find_major_hr_declines(sm)
        t = tick("find_major_hr_declines", t)
        frontend.calculate_hrr(sm)
        t = tick("calculate_hrr", t)
        frontend.find_peak_recovery_rate(sm)
        t = tick("find_peak_recovery_rate", t)
        frontend.find_peak_exertion_rate(sm)
        t = tick("find_peak_exertion_rate", t)
        frontend.calculate_windowed_hrv(bi, r, p)
        t = tick("calculate_windowed_hrv", t)
    out = {k: round(v / reps * 1e3, 4) for k, v in acc.items()}
    out["total_ms"] = round(sum(out.values()), 4)
    # inside preprocess: host gather alone
    import ctypes as C
    from bpm_analysis_b200 import classifier
    lib = classifier.load_host_library()
    stage = torch.empty(len(env), dtype=torch.int16).pin_memory()
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.bpm_host_gather_frames(C.c_void_p(pcm.ctypes.data), 2, len(pcm), 159, C.c_void_p(stage.data_ptr()), 0)
    out["host_gather_alone_ms"] = round((time.perf_counter() - t0) / reps * 1e3, 4)
    out["host_threads"] = int(lib.bpm_host_threads())
    out["cpu_count"] = os.cpu_count()
    out["trace_ms_per_step"] = {k: round(v / reps * 1e3, 4) for k, v in svc.trace.items()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
