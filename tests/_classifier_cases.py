"""Seeded inputs for the classifier parity tests: the state ``PeakClassifier.__init__`` hands to
``classify_peaks`` (envelope, noise floor, raw peaks, smoothed deviation series, params, start
BPM, recovery window), built directly so that thousands of decisions run in milliseconds.

Used by ``oracle/make_golden_classifier.py`` (reference outputs -> tests/golden) and by
``tests/test_classifier_cpu.py`` (same inputs -> libbpm_host.so).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from bpm_analysis_b200.params import default_params

RATE = 300


def make_case(seed: int):
    """-> dict(env, floor, peaks, dev_index, dev_values, params, start_bpm, peak_time, recovery_time, rate)."""
    rng = np.random.default_rng(seed)
    kind = seed % 6
    bpm0 = float(rng.uniform(45, 190))
    n_beats = int(rng.integers(30, 400))
    t, times, amps = 0.5, [], []
    bpm = bpm0
    for _ in range(n_beats):
        bpm = float(np.clip(bpm + rng.normal(0, 1.5), 38, 230))
        rr = 60.0 / bpm * (1 + 0.03 * rng.standard_normal())
        s1 = float(rng.uniform(0.5, 1.5))
        times.append(t); amps.append(s1)
        if rng.random() < (0.9 if kind in (0, 1) else 0.5):                 # S2
            times.append(t + min(0.30, 0.35 * rr) * (1 + 0.05 * rng.standard_normal()))
            amps.append(s1 * float(rng.uniform(0.2, 1.8 if kind == 2 else 0.9)))
        if rng.random() < (0.02, 0.1, 0.3, 0.6, 0.05, 0.9)[kind]:           # noise peak in diastole
            times.append(t + rr * float(rng.uniform(0.5, 0.9)))
            amps.append(float(rng.uniform(0.01, 1.2)))
        if kind == 4 and rng.random() < 0.1:                                # dropped stretch
            t += rr * float(rng.integers(2, 6))
        t += rr
    order = np.argsort(times)
    idx = np.round(np.asarray(times)[order] * RATE).astype(np.int64)
    keep = np.concatenate([[True], np.diff(idx) > 0])
    peaks = idx[keep]
    amps = np.asarray(amps)[order][keep]
    m = int(peaks[-1]) + RATE
    env = np.abs(rng.normal(0.02, 0.01, m))
    env[peaks] = amps
    floor = np.abs(rng.normal(0.03, 0.01, m)) * (5.0 if kind == 3 else 1.0)
    if kind == 5:
        floor[peaks[::7]] = env[peaks[::7]] + 0.1                           # zero-strength peaks

    params = default_params()
    if seed % 4 == 1:
        params.update(pairing_confidence_threshold=float(rng.uniform(0.3, 0.8)),
                      lone_s1_confidence_threshold=float(rng.uniform(0.3, 0.8)),
                      stability_history_window=int(rng.integers(3, 30)),
                      kickstart_check_threshold=float(rng.uniform(0.1, 0.9)),
                      cascade_reset_trigger_count=int(rng.integers(1, 5)),
                      s1_s2_interval_cap_sec=float(rng.uniform(0.15, 0.5)),
                      interval_penalty_start_factor=float(rng.uniform(0.8, 1.2)),
                      min_bpm=int(rng.integers(30, 60)), max_bpm=int(rng.integers(150, 260)))
    if seed % 4 == 2:
        params.update(enable_interval_penalty=False, contractility_bpm_low=float(rng.uniform(60, 110)),
                      contractility_bpm_high=float(rng.uniform(120, 180)), s1_s2_boost_ratio=float(rng.uniform(1.0, 2.0)))
    # the arithmetic of PeakClassifier._initialize_state (bpm_analysis.py:92-100) with pandas
    strength = env[peaks] - floor[peaks]
    strength[strength < 0] = 0
    dev = np.abs(np.diff(strength)) / (np.maximum(strength[:-1], strength[1:]) + 1e-9)
    dev_index = (peaks[:-1] + peaks[1:]) / 2 / RATE
    window = max(5, int(len(dev) * params["deviation_smoothing_factor"]))
    smoothed = pd.Series(dev, index=dev_index).rolling(window=window, min_periods=1, center=True).mean()
    total = float(peaks[-1]) / RATE
    has_window = seed % 3 != 0
    return {"env": env, "floor": floor, "peaks": peaks, "dev_index": smoothed.index.values.copy(),
            "dev_values": smoothed.values.copy(), "params": params,
            "start_bpm": None if seed % 5 == 0 else float(np.round(bpm0 * rng.uniform(0.7, 1.3), 1)),
            "peak_time": float(total * 0.3) if has_window else None,
            "recovery_time": float(total * 0.7) if has_window else None, "rate": RATE}


class ClassifierStandIn:
    """The attributes ``classify_peaks`` reads from a ``PeakClassifier`` (bpm_analysis.py:71-111)."""

    def __init__(self, case):
        self.audio_envelope = case["env"]
        self.sample_rate = case["rate"]
        self.params = case["params"]
        self.peak_bpm_time_sec = case["peak_time"]
        self.recovery_end_time_sec = case["recovery_time"]
        self.state = state_of(case)


def state_of(case):
    floor = pd.Series(case["floor"], index=np.arange(len(case["floor"])))
    dev = pd.Series(case["dev_values"], index=case["dev_index"])
    hint = case["start_bpm"]
    return {"analysis_data": {"dynamic_noise_floor_series": floor, "trough_indices": np.array([], dtype=np.int64),
                              "deviation_series": dev},
            "dynamic_noise_floor": floor, "trough_indices": np.array([], dtype=np.int64), "all_peaks": case["peaks"],
            "smoothed_dev_series": dev, "long_term_bpm": float(hint) if hint else 80.0, "candidate_beats": [],
            "beat_debug_info": {}, "long_term_bpm_history": [], "sorted_troughs": [], "consecutive_rr_rejections": 0,
            "loop_idx": 0}


def pack_result(result):
    """(final_peaks, all_peaks, analysis_data) -> plain containers for comparison / JSON."""
    final_peaks, all_peaks, data = result
    series = data.get("long_term_bpm_series")
    return {"final_peaks": [int(x) for x in final_peaks],
            "keys": [int(k) for k in data["beat_debug_info"].keys()],
            "texts": list(data["beat_debug_info"].values()),
            "lt_times": [] if series is None else [float(x).hex() for x in series.index.values],
            "lt_values": [] if series is None else [float(x).hex() for x in series.values]}
