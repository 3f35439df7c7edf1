#!/usr/bin/env python
"""Time the whole drop-in chain for one recording, stage by stage, the way `analyze_wav_file`
(bpm_analysis.py:1725-1768) strings the pieces together -- minus plotting and report writing:

  a1 preprocess -> a2 noise floor -> preliminary classifier pass (threshold 0.75) -> start BPM /
  recovery window -> main classifier pass -> correction passes -> BPM series, slopes, HRV.

    python tools/chain_probe.py [duration_sec] [sample_rate]

The GPU front end (frontend.py -> libbpm_b200.so) + the compiled sequential stage
(libbpm_host.so).  Prints one JSON line.  The same chain with the CPU oracle as front end (the
"before" column, and a dry run on a machine without a GPU) is tests/chain_probe_cpu_oracle.py,
which reuses run() below -- only tests/ may import oracle/.
"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from bpm_analysis_b200 import classifier, corrections, synth          # noqa: E402
from bpm_analysis_b200.params import default_params                   # noqa: E402


class Clf:
    """The attributes of the reference's PeakClassifier that the drop-ins read."""

    def __init__(self, env, rate, params, hint, floor, troughs, peak_t, rec_t, init_state):
        self.audio_envelope, self.sample_rate, self.params = env, rate, params
        self.peak_bpm_time_sec, self.recovery_end_time_sec = peak_t, rec_t
        self.state = init_state(self, hint, floor, troughs)


class GpuFrontEnd:
    """The drop-in functions of frontend.py, bound to one params dict."""

    label = "gpu"

    def __init__(self, sr, params):
        import torch
        from bpm_analysis_b200 import frontend as fe
        torch.cuda.set_device(0)
        self.fe, self.sr, self.params = fe, sr, params
        self.init_state = fe._initialize_state
        self.metrics = [fe.find_major_hr_inclines, fe.find_major_hr_declines, fe.find_peak_recovery_rate,
                        fe.find_peak_exertion_rate]

    def preprocess(self, x):
        env, rate, _, _ = self.fe.preprocess_pcm(x, self.sr, self.params)
        return env, rate

    def noise_floor(self, env, rate):
        return self.fe._calculate_dynamic_noise_floor(env, rate, self.params)

    def bpm_series(self, beats, rate):
        return self.fe.calculate_bpm_series(beats, rate, self.params)

    def hrv(self, beats, rate):
        return self.fe.calculate_windowed_hrv(beats, rate, self.params)


def parse_args():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    dur = float(args[0]) if args else 3600.0
    sr = int(args[1]) if len(args) > 1 else 48000
    return dur, sr


def run(front_end_factory):
    dur, sr = parse_args()
    params = default_params()
    params["save_filtered_wav"] = False
    pcm, sr, _ = synth.config_c2(seed=2, duration_sec=dur) if sr == 48000 else synth.pcg_recording(dur, sr, lambda t: 75.0, 2)
    F = front_end_factory(sr, params)
    preprocess, noise_floor, init_state = F.preprocess, F.noise_floor, F.init_state
    bpm_series, metrics, hrv = F.bpm_series, F.metrics, F.hrv

    def chain():
        t = {}
        c0 = time.perf_counter()
        env, rate = preprocess(pcm)
        t["a1_preprocess"] = time.perf_counter() - c0
        c = time.perf_counter()
        floor, troughs = noise_floor(env, rate)
        t["a2_noise_floor"] = time.perf_counter() - c
        c = time.perf_counter()
        p1 = dict(params, pairing_confidence_threshold=0.75)                     # _run_preliminary_pass, :1622-1653
        pre = Clf(env, rate, p1, None, floor, troughs, None, None, init_state)
        t["a3_a4_prelim_state"] = time.perf_counter() - c
        c = time.perf_counter()
        anchors, _, _ = classifier.classify_peaks(pre)
        t["classifier_prelim"] = time.perf_counter() - c
        c = time.perf_counter()
        start_bpm = 80.0
        if len(anchors) >= 10:
            med = np.median(np.diff(anchors) / rate)
            if med > 0:
                start_bpm = 60.0 / med
        series, times = bpm_series(anchors, rate)
        peak_t = rec_t = None
        if times is not None and len(times) >= 2:                                 # find_recovery_phase, :1612-1620
            peak_t = times[np.argmax(series.to_numpy())]
            rec_t = peak_t + params.get("recovery_phase_duration_sec", 120.0)
        t["prelim_bpm_series"] = time.perf_counter() - c
        c = time.perf_counter()
        main_clf = Clf(env, rate, params, start_bpm, floor, troughs, peak_t, rec_t, init_state)
        t["a3_a4_main_state"] = time.perf_counter() - c
        c = time.perf_counter()
        s1, raw, data = classifier.classify_peaks(main_clf)
        t["classifier_main"] = time.perf_counter() - c
        c = time.perf_counter()
        peaks = corrections.correct_peaks_by_rhythm(s1, env, rate, params)
        info = data["beat_debug_info"]
        for _ in range(5):
            peaks, info, made = corrections._fix_rhythmic_discontinuities(peaks, raw, info, env, floor, params, rate)
            if made == 0:
                break
        t["correction_passes"] = time.perf_counter() - c
        c = time.perf_counter()
        series, _ = bpm_series(peaks, rate)
        for fn in metrics:
            fn(series)
        hrv(peaks, rate)
        t["a5_a8_beat_metrics"] = time.perf_counter() - c
        t["total"] = time.perf_counter() - c0
        return t, {"M": int(len(env)), "raw_peaks": int(len(raw)), "anchors": int(len(anchors)), "beats": int(len(peaks))}

    import logging
    logging.getLogger().setLevel(logging.WARNING)
    chain()                                                                       # warm-up (allocations, library loads)
    runs = [chain() for _ in range(3)]
    best = min(runs, key=lambda r: r[0]["total"])
    print(json.dumps({"front_end": F.label, "duration_sec": dur, "sample_rate": sr,
                      "counts": best[1], "ms": {k: round(1e3 * v, 2) for k, v in best[0].items()},
                      "audio_hours_per_sec_whole_chain": dur / 3600.0 / best[0]["total"]}))


if __name__ == "__main__":
    run(GpuFrontEnd)
