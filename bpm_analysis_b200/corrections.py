"""Drop-in replacements for the reference's correction passes:
``correct_peaks_by_rhythm`` (bpm_analysis.py:1257-1306) and ``_fix_rhythmic_discontinuities``
(:1309-1412), with their per-beat Python loops in compiled host code (``csrc/corrections.cpp`` ->
``libbpm_host.so``, ``include/bpm_host.h``).

The vectorised numpy calls that produce the thresholds (``np.diff``, ``np.median``,
``np.percentile``) are the reference's; the compiled loops return their decisions as events, which
are applied to the debug-string dict and logged here in the reference's words -- same peaks, same
dict, same log lines, same correction count.
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


def _say(text: str) -> None:
    logging.info(text)


def _dbg(text: str) -> None:
    logging.info("[Correction DEBUG] " + text)


def _sec(sample, rate) -> str:
    return format(sample / rate, ".2f")


def _rr(beats, rate) -> np.ndarray:
    """Intervals between consecutive beats in seconds (the reference's ``np.diff(...) / sample_rate``)."""
    return np.diff(beats) / rate


def correct_peaks_by_rhythm(peaks: np.ndarray, audio_envelope: np.ndarray, sample_rate: int, params: Dict) -> np.ndarray:
    """Mirrors bpm_analysis.py:1257-1306: a beat closer to its predecessor than a fraction of the median
    interval is in conflict with it, the one with the larger envelope amplitude stays."""
    n_in = len(peaks)
    if n_in < 5:
        return peaks
    _say(f"--- STAGE 4: Correcting peaks based on rhythm. Initial count: {n_in} ---")
    med = np.median(_rr(peaks, sample_rate))
    limit = med * params.get("rr_correction_threshold_pct", 0.6)
    _say("Median R-R: {:.3f}s. Correction threshold: {:.3f}s.".format(med, limit))
    pk = np.ascontiguousarray(peaks, dtype=np.int64)
    env = np.ascontiguousarray(audio_envelope, dtype=np.float64)
    out = np.empty(n_in, dtype=np.int64)
    events = (CorrectionEvent * n_in)()
    n_out, n_ev = C.c_int64(), C.c_int64()
    rc = _lib().bpm_correct_peaks_by_rhythm(pk.ctypes.data, n_in, env.ctypes.data, len(env), float(sample_rate),
                                            float(limit), out.ctypes.data, C.byref(n_out), events, C.byref(n_ev))
    if rc != 0:
        raise ValueError(f"bpm_correct_peaks_by_rhythm rejected its arguments (code {rc})")
    for e in events[:n_ev.value]:
        if e.kind == 1:
            _say(f"Conflict at {_sec(e.a, sample_rate)}s. Replaced previous peak at {_sec(e.b, sample_rate)}s "
                 "due to higher amplitude.")
        else:
            _say(f"Conflict at {_sec(e.a, sample_rate)}s. Discarding current peak due to lower amplitude.")
    kept = n_out.value
    if kept < n_in:
        _say(f"Correction complete. Removed {n_in - kept} peak(s). Final count: {kept}")
    else:
        _say("Correction pass complete. No rhythmic conflicts found.")
    return out[:kept].copy()


def _discontinuity_thresholds(s1_peaks, sample_rate, params):
    """(median interval, short limit, long limit) of bpm_analysis.py:1319-1330 -- the median is taken over the
    intervals inside Tukey's fences -- or None when no interval is."""
    rr = _rr(s1_peaks, sample_rate)
    lo_q, hi_q = np.percentile(rr, [25, 75])
    spread = hi_q - lo_q
    inside = rr[(rr > (lo_q - 1.5 * spread)) & (rr < (hi_q + 1.5 * spread))]
    if len(inside) < 1:
        return None
    med = np.median(inside)
    return med, med * params["rr_correction_threshold_pct"], med * params.get("rr_correction_long_interval_pct", 1.7)


def _fix_rhythmic_discontinuities(s1_peaks: np.ndarray, all_raw_peaks: np.ndarray, debug_info: Dict,
                                  audio_envelope: np.ndarray, dynamic_noise_floor: pd.Series, params: Dict,
                                  sample_rate: int) -> Tuple[np.ndarray, Dict, int]:
    """Mirrors bpm_analysis.py:1309-1412."""
    margin = 3
    n_s1 = len(s1_peaks)
    if n_s1 < margin * 2:
        _dbg(f"Skipping correction pass: Not enough S1 peaks ({n_s1}) to apply a margin of {margin}.")
        return s1_peaks, debug_info, 0
    limits = _discontinuity_thresholds(s1_peaks, sample_rate, params)
    if limits is None:
        _dbg("Not enough stable R-R intervals to determine median. Skipping correction.")
        return s1_peaks, debug_info, 0
    med, short_limit, long_limit = limits
    _dbg("Median R-R: {:.3f}s. Short Threshold: < {:.3f}s. Long Threshold: > {:.3f}s.".format(med, short_limit, long_limit))
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
        len(env), float(sample_rate), float(short_limit), float(long_limit), float(waiver_strength), float(waiver_ratio),
        out.ctypes.data, C.byref(n_out), events, cap, C.byref(n_ev), C.byref(n_corr))
    if rc != 0:
        raise ValueError(f"bpm_fix_rhythmic_discontinuities rejected its arguments (code {rc})")

    relabelled = dict(debug_info)
    _dbg(f"Checking for long intervals between beat {margin} and beat {n_s1 - margin}...")
    short_phase = False
    for e in events[:n_ev.value]:
        if e.kind >= 5 and not short_phase:
            _dbg("Starting SHORT interval check...")
            short_phase = True
        if e.kind == 3:
            _dbg(f"Found LONG interval at {_sec(e.a, sample_rate)}s. Investigating gap...")
        elif e.kind == 4:
            first, second = raw_keys[e.a], raw_keys[e.b]
            _dbg(f"  - SUCCESS: Re-labeling S1/S2 pair at {_sec(first, sample_rate)}s.")
            for key, label in ((first, S1_CORRECTED_GAP), (second, S2_CORRECTED_GAP)):
                relabelled[key] = f"{label}§ORIGINAL_REASON§{relabelled.get(key, 'Noise')}"
        elif e.kind == 5:
            _dbg(f"Found SHORT interval of {e.x:.3f}s between beats at {_sec(e.a, sample_rate)}s and "
                 f"{_sec(e.b, sample_rate)}s. Resolving...")
        elif e.kind == 6:
            _dbg(f"  - Removing weaker peak at {_sec(e.a, sample_rate)}s.")
    if not short_phase:
        _dbg("Starting SHORT interval check...")
    return out[:n_out.value].copy(), relabelled, int(n_corr.value)


def install(ref_module):
    """Rebind the two correction passes on an imported reference module (resolved from its globals
    by ``_refine_and_correct_peaks``, bpm_analysis.py:1664-1680)."""
    ref_module.correct_peaks_by_rhythm = correct_peaks_by_rhythm
    ref_module._fix_rhythmic_discontinuities = _fix_rhythmic_discontinuities
    return ref_module
