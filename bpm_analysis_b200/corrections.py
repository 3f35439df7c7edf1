"""Drop-in replacements for the reference's correction passes:
``correct_peaks_by_rhythm`` (bpm_analysis.py:1257-1306) and ``_fix_rhythmic_discontinuities``
(:1309-1412), with their per-beat Python loops in compiled host code (``csrc/corrections.cpp`` ->
``libbpm_host.so``, ``include/bpm_host.h``).

The vectorised numpy calls that produce the thresholds (``np.diff``, ``np.median``,
``np.percentile``) are made exactly as the reference makes them; the compiled loops return their
decisions as events, which are applied to the debug-string dict and logged here with the
reference's own f-strings -- same peaks, same dict, same log lines, same correction count.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import Dict, Tuple

import numpy as np
import pandas as pd

from .classifier import load_host_library

S1_CORRECTED_GAP = "S1 (Paired - Corrected from Gap)"       # PeakType values, bpm_analysis.py:34-35
S2_CORRECTED_GAP = "S2 (Paired - Corrected from Gap)"

EXPORTED_SYMBOLS = ("bpm_correct_peaks_by_rhythm", "bpm_fix_rhythmic_discontinuities")


class CorrectionEvent(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("a", C.c_int64), ("b", C.c_int64), ("x", C.c_double)]


_bound = False


def _lib():
    global _bound
    lib = load_host_library()
    if not _bound:
        ev = C.POINTER(CorrectionEvent)
        i64p = C.POINTER(C.c_int64)
        lib.bpm_correct_peaks_by_rhythm.restype = C.c_int
        lib.bpm_correct_peaks_by_rhythm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                                    C.c_void_p, i64p, ev, i64p]
        lib.bpm_fix_rhythmic_discontinuities.restype = C.c_int
        lib.bpm_fix_rhythmic_discontinuities.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                                         C.c_double, C.c_double, C.c_double, C.c_void_p, i64p, ev,
                                                         C.c_int64, i64p, i64p]
        _bound = True
    return lib


def correct_peaks_by_rhythm(peaks: np.ndarray, audio_envelope: np.ndarray, sample_rate: int, params: Dict) -> np.ndarray:
    """Mirrors bpm_analysis.py:1257-1306."""
    if len(peaks) < 5:
        return peaks
    logging.info(f"--- STAGE 4: Correcting peaks based on rhythm. Initial count: {len(peaks)} ---")
    rr_intervals_sec = np.diff(peaks) / sample_rate
    median_rr_sec = np.median(rr_intervals_sec)
    correction_threshold_sec = median_rr_sec * params.get("rr_correction_threshold_pct", 0.6)
    logging.info(f"Median R-R: {median_rr_sec:.3f}s. Correction threshold: {correction_threshold_sec:.3f}s.")
    pk = np.ascontiguousarray(peaks, dtype=np.int64)
    env = np.ascontiguousarray(audio_envelope, dtype=np.float64)
    out = np.empty(len(pk), dtype=np.int64)
    events = (CorrectionEvent * len(pk))()
    n_out, n_ev = C.c_int64(), C.c_int64()
    rc = _lib().bpm_correct_peaks_by_rhythm(pk.ctypes.data, len(pk), env.ctypes.data, len(env), float(sample_rate),
                                            float(correction_threshold_sec), out.ctypes.data, C.byref(n_out), events,
                                            C.byref(n_ev))
    if rc != 0:
        raise ValueError(f"bpm_correct_peaks_by_rhythm rejected its arguments (code {rc})")
    for i in range(n_ev.value):
        e = events[i]
        if e.kind == 1:
            logging.info(f"Conflict at {e.a/sample_rate:.2f}s. Replaced previous peak at {e.b/sample_rate:.2f}s due to higher amplitude.")
        else:
            logging.info(f"Conflict at {e.a/sample_rate:.2f}s. Discarding current peak due to lower amplitude.")
    final_peak_count = n_out.value
    if final_peak_count < len(peaks):
        logging.info(f"Correction complete. Removed {len(peaks) - final_peak_count} peak(s). Final count: {final_peak_count}")
    else:
        logging.info("Correction pass complete. No rhythmic conflicts found.")
    return out[:final_peak_count].copy()


def _fix_rhythmic_discontinuities(s1_peaks: np.ndarray, all_raw_peaks: np.ndarray, debug_info: Dict,
                                  audio_envelope: np.ndarray, dynamic_noise_floor: pd.Series, params: Dict,
                                  sample_rate: int) -> Tuple[np.ndarray, Dict, int]:
    """Mirrors bpm_analysis.py:1309-1412."""
    def log_debug(msg):
        logging.info(f"[Correction DEBUG] {msg}")

    margin = 3
    if len(s1_peaks) < margin * 2:
        log_debug(f"Skipping correction pass: Not enough S1 peaks ({len(s1_peaks)}) to apply a margin of {margin}.")
        return s1_peaks, debug_info, 0
    rr_intervals_sec = np.diff(s1_peaks) / sample_rate
    q1, q3 = np.percentile(rr_intervals_sec, [25, 75])
    iqr = q3 - q1
    stable_rr_intervals = rr_intervals_sec[
        (rr_intervals_sec > (q1 - 1.5 * iqr)) & (rr_intervals_sec < (q3 + 1.5 * iqr))]
    if len(stable_rr_intervals) < 1:
        log_debug("Not enough stable R-R intervals to determine median. Skipping correction.")
        return s1_peaks, debug_info, 0
    median_rr_sec = np.median(stable_rr_intervals)
    short_conflict_threshold_sec = median_rr_sec * params["rr_correction_threshold_pct"]
    long_conflict_threshold_sec = median_rr_sec * params.get("rr_correction_long_interval_pct", 1.7)
    log_debug(
        f"Median R-R: {median_rr_sec:.3f}s. Short Threshold: < {short_conflict_threshold_sec:.3f}s. Long Threshold: > {long_conflict_threshold_sec:.3f}s.")
    waiver_strength, waiver_ratio = params["penalty_waiver_strength_ratio"], params["penalty_waiver_max_s2_s1_ratio"]

    s1 = np.ascontiguousarray(s1_peaks, dtype=np.int64)
    raw = np.ascontiguousarray(all_raw_peaks, dtype=np.int64)
    raw_keys = list(all_raw_peaks)
    is_noise = np.fromiter(("Noise" in debug_info.get(p, "") for p in raw_keys), dtype=np.uint8, count=len(raw_keys))
    env = np.ascontiguousarray(audio_envelope, dtype=np.float64)
    floor = np.ascontiguousarray(dynamic_noise_floor.values, dtype=np.float64)
    out = np.empty(len(s1) + len(raw), dtype=np.int64)
    cap = 4 * (len(s1) + len(raw)) + 8
    events = (CorrectionEvent * cap)()
    n_out, n_ev, n_corr = C.c_int64(), C.c_int64(), C.c_int64()
    rc = _lib().bpm_fix_rhythmic_discontinuities(
        s1.ctypes.data, len(s1), raw.ctypes.data, len(raw), is_noise.ctypes.data, env.ctypes.data, floor.ctypes.data,
        len(env), float(sample_rate), float(short_conflict_threshold_sec), float(long_conflict_threshold_sec),
        float(waiver_strength), float(waiver_ratio), out.ctypes.data, C.byref(n_out), events, cap, C.byref(n_ev),
        C.byref(n_corr))
    if rc != 0:
        raise ValueError(f"bpm_fix_rhythmic_discontinuities rejected its arguments (code {rc})")

    corrected_debug_info = debug_info.copy()
    log_debug(f"Checking for long intervals between beat {margin} and beat {len(s1_peaks) - margin}...")
    short_started = False
    for i in range(n_ev.value):
        e = events[i]
        if e.kind >= 5 and not short_started:
            log_debug("Starting SHORT interval check...")
            short_started = True
        if e.kind == 3:
            log_debug(f"Found LONG interval at {e.a / sample_rate:.2f}s. Investigating gap...")
        elif e.kind == 4:
            candidate_s1, candidate_s2 = raw_keys[e.a], raw_keys[e.b]
            log_debug(f"  - SUCCESS: Re-labeling S1/S2 pair at {candidate_s1 / sample_rate:.2f}s.")
            original_reason_s1 = corrected_debug_info.get(candidate_s1, "Noise")
            corrected_debug_info[candidate_s1] = f"{S1_CORRECTED_GAP}§ORIGINAL_REASON§{original_reason_s1}"
            original_reason_s2 = corrected_debug_info.get(candidate_s2, "Noise")
            corrected_debug_info[candidate_s2] = f"{S2_CORRECTED_GAP}§ORIGINAL_REASON§{original_reason_s2}"
        elif e.kind == 5:
            log_debug(
                f"Found SHORT interval of {e.x:.3f}s between beats at {e.a / sample_rate:.2f}s and {e.b / sample_rate:.2f}s. Resolving...")
        elif e.kind == 6:
            log_debug(f"  - Removing weaker peak at {e.a / sample_rate:.2f}s.")
    if not short_started:
        log_debug("Starting SHORT interval check...")
    return out[:n_out.value].copy(), corrected_debug_info, int(n_corr.value)


def install(ref_module):
    """Rebind the two correction passes on an imported reference module (resolved from its globals
    by ``_refine_and_correct_peaks``, bpm_analysis.py:1664-1680)."""
    ref_module.correct_peaks_by_rhythm = correct_peaks_by_rhythm
    ref_module._fix_rhythmic_discontinuities = _fix_rhythmic_discontinuities
    return ref_module
