"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module, and there only
as the checker / the timed CPU comparator.  ``bpm_analysis_b200`` itself must
fail loudly when its CUDA library is missing; it never routes through here.

What this is: a restatement, array-in / array-out, of the data-parallel front
end of the reference's ``bpm_analysis.py`` (pixeru/bpm_analysis).  The
reference has no native code of its own: every float operation on this path is
a call into numpy / scipy.signal / pandas (un-vendored and unpinned in the
reference: ``hugging-face-space/requirements.txt:1-7``).  The versions this
oracle was pinned with are the ones in this image: numpy 2.3.5, scipy 1.18.1,
pandas 3.0.2.  The restatement therefore calls the same library entry points
with the same arguments, in the same order, as the reference lines cited on
each function.

Parity pin: ``tests/test_oracle_golden.py`` checks every function here against
(i) golden vectors produced by importing the unmodified reference module in
the authoring container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``)
and (ii) the reference's own shipped run ``samples/vulpine_*`` (raw-peak and
trough indices, BPM-series CSV, HRV / slope summary), re-encoded in
``tests/golden/vulpine.npz``.  ``fullrate`` filter mode has no reference
behaviour (the reference decimates first); for that mode the oracle is
``scipy.signal.sosfiltfilt`` at the original rate followed by the reference's
own envelope lines -- "parity unpinned" by any reference artefact.
"""
from __future__ import annotations

import datetime
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
from scipy.signal import butter, filtfilt, find_peaks, sosfiltfilt

LOWCUT_HZ, HIGHCUT_HZ = 20.0, 150.0       # bpm_analysis.py:1018


# --------------------------------------------------------------------------- a1
def decimation_plan(sample_rate: int, params: Dict) -> Tuple[int, int]:
    """(ds, envelope_rate).  bpm_analysis.py:1021-1036."""
    ds = params["downsample_factor"]
    highcut = float(params.get("highcut_hz", HIGHCUT_HZ))
    max_safe = int((sample_rate / (highcut * 2)) - 1)
    if ds > max_safe:
        ds = max(1, max_safe)
    if ds > 1:
        return int(ds), int(sample_rate // ds)
    return 1, int(sample_rate)


def bandpass_filtered(audio_data: np.ndarray, sample_rate: int, params: Dict
                      ) -> Tuple[np.ndarray, int]:
    """Zero-phase band-passed signal at the envelope rate.  bpm_analysis.py:1015-1045.

    ``filter_mode == 'parity'`` (default) is the reference: ``x[::ds]`` then
    ``filtfilt(butter(2, [lo, hi], 'band'), x)``.  ``'fullrate'`` filters at
    ``sample_rate`` with ``sosfiltfilt`` (same 15-sample odd padding, SURVEY
    Appendix A.1) and decimates afterwards.
    """
    if audio_data.ndim > 1:
        audio_data = np.mean(audio_data, axis=1)                    # :1016
    lowcut = float(params.get("lowcut_hz", LOWCUT_HZ))
    highcut = float(params.get("highcut_hz", HIGHCUT_HZ))
    ds, nsr = decimation_plan(sample_rate, params)
    mode = params.get("filter_mode", "parity")
    if mode == "parity":
        x = audio_data[::ds] if ds > 1 else audio_data               # :1033
        nyq = 0.5 * nsr
        lo, hi = lowcut / nyq, highcut / nyq
        if hi >= 1.0:                                                # :1041
            raise ValueError(f"Cannot create a {highcut}Hz filter. The effective sample "
                             f"rate of {nsr}Hz is too low.")
        b, a = butter(2, [lo, hi], btype="band")                     # :1044
        return filtfilt(b, a, x), nsr                                # :1045
    nyq = 0.5 * sample_rate
    sos = butter(2, [lowcut / nyq, highcut / nyq], btype="band", output="sos")
    y = sosfiltfilt(sos, np.asarray(audio_data, dtype=np.float64), padlen=15)
    return (y[::ds] if ds > 1 else y), nsr


def envelope_of(filtered: np.ndarray, envelope_rate: int) -> np.ndarray:
    """|y| then centred rolling mean, window rate//10, min_periods=1.  :1052-1054."""
    w = envelope_rate // 10
    return pd.Series(np.abs(filtered)).rolling(window=w, min_periods=1, center=True).mean().values


def debug_wav_samples(filtered: np.ndarray) -> np.ndarray:
    """int16 side output (truncating cast).  :1047-1049, :1056-1059."""
    return np.int16(filtered / np.max(np.abs(filtered)) * 32767)


def preprocess_pcm(audio_data: np.ndarray, sample_rate: int, params: Dict
                   ) -> Tuple[np.ndarray, int, np.ndarray]:
    """Array-level body of ``preprocess_audio`` (:1007-1062): (envelope, rate, filtered)."""
    y, nsr = bandpass_filtered(audio_data, sample_rate, params)
    return envelope_of(y, nsr), nsr, y


# --------------------------------------------------------------------------- a2
def _interp_troughs(envelope: np.ndarray, troughs: np.ndarray) -> pd.Series:
    s = pd.Series(index=troughs, data=envelope[troughs])              # :1081 / :1103
    return s.reindex(np.arange(len(envelope))).interpolate()         # :1082 / :1104


def calculate_dynamic_noise_floor(envelope: np.ndarray, rate: int, params: Dict
                                  ) -> Tuple[pd.Series, np.ndarray]:
    """bpm_analysis.py:1064-1117."""
    dist = int(params["min_peak_distance_sec"] * rate)                # :1066
    prom = np.quantile(envelope, params["trough_prominence_quantile"])  # :1067
    all_troughs, _ = find_peaks(-envelope, distance=dist, prominence=prom)  # :1070
    n = len(envelope)
    if len(all_troughs) < 5:                                          # :1073-1077
        fb = np.quantile(envelope, params["noise_floor_quantile"])
        return pd.Series(fb, index=np.arange(n)), all_troughs
    w = int(params["noise_window_sec"] * rate)                        # :1083
    q = params["noise_floor_quantile"]
    draft = _interp_troughs(envelope, all_troughs).rolling(
        window=w, min_periods=3, center=True).quantile(q)             # :1085
    draft = draft.bfill().ffill()                                     # :1086
    mult = params.get("trough_rejection_multiplier", 4.0)             # :1091
    kept: List[int] = []
    dv = draft.values
    for t in all_troughs:                                             # :1092-1097
        f = dv[t]
        if not np.isnan(f) and envelope[t] <= mult * f:
            kept.append(t)
    if len(kept) > 2:                                                 # :1102-1106
        floor = _interp_troughs(envelope, np.asarray(kept)).rolling(
            window=w, min_periods=3, center=True).quantile(q)
        floor = floor.bfill().ffill()
    else:                                                             # :1107-1110
        floor = draft
    if floor.isnull().all():                                          # :1113-1115
        floor = pd.Series(np.quantile(envelope, 0.1), index=np.arange(n))
    return floor, np.array(kept)


# --------------------------------------------------------------------------- a3
def find_raw_peaks(envelope: np.ndarray, rate: int, params: Dict,
                   height_threshold: np.ndarray) -> np.ndarray:
    """PeakClassifier._find_raw_peaks, bpm_analysis.py:223-229."""
    prom = np.quantile(envelope, params["peak_prominence_quantile"])
    dist = int(params["min_peak_distance_sec"] * rate)
    peaks, _ = find_peaks(envelope, height=height_threshold, prominence=prom, distance=dist)
    return peaks


# --------------------------------------------------------------------------- a4
def peak_metrics(envelope: np.ndarray, rate: int, params: Dict, floor: pd.Series,
                 peaks: np.ndarray) -> Dict[str, object]:
    """Numeric part of PeakClassifier._initialize_state, bpm_analysis.py:93-100."""
    floor_at = floor.reindex(peaks, method="nearest").values          # :93
    strength = envelope[peaks] - floor_at                             # :94
    strength[strength < 0] = 0                                        # :95
    dev = np.abs(np.diff(strength)) / (np.maximum(strength[:-1], strength[1:]) + 1e-9)  # :96
    times = (peaks[:-1] + peaks[1:]) / 2 / rate                       # :97
    series = pd.Series(dev, index=times)                              # :98
    win = max(5, int(len(series) * params["deviation_smoothing_factor"]))  # :99
    smoothed = series.rolling(window=win, min_periods=1, center=True).mean()  # :100
    return {"strength": strength, "deviation": dev, "times": times,
            "smoothed_dev_series": smoothed, "window": win}


def peak_trough_noise(envelope: np.ndarray, floor: np.ndarray, peaks: np.ndarray, troughs: np.ndarray,
                      params: Dict) -> Dict[str, np.ndarray]:
    """Surrounding-trough noise per raw peak.  PARITY UNPINNED: bpm_analysis.py no longer has a
    function for it; restated from Documentation/Changelog.md:454 ("the deeper of the two
    troughs surrounding a peak"), "BPM Detection logic explained.md":262 (trough amplitude >
    3 x noise floor => noisy; config.py:31) and :276-278 (look-ahead veto
    2*(peak - next_trough) < (next_peak - next_trough); config.py:30 made the 2 a parameter)."""
    nm = float(params.get("trough_noise_multiplier", 3.0))
    vm = float(params.get("trough_veto_multiplier", 2.1))
    peaks = np.asarray(peaks, dtype=np.int64)
    troughs = np.asarray(troughs, dtype=np.int64)
    n = len(peaks)
    prev = np.full(n, np.nan)
    nxt = np.full(n, np.nan)
    ratio = np.full(n, np.nan)
    flags = np.zeros(n, dtype=np.uint8)
    for k, p in enumerate(peaks):
        i_next = int(np.searchsorted(troughs, p, side="right"))
        i_prev = int(np.searchsorted(troughs, p, side="left")) - 1
        if i_prev >= 0:
            prev[k] = envelope[troughs[i_prev]]
        if i_next < len(troughs):
            nxt[k] = envelope[troughs[i_next]]
        deeper = np.nanmin([prev[k], nxt[k]]) if (i_prev >= 0 or i_next < len(troughs)) else np.nan
        ratio[k] = deeper / floor[p]
        if ratio[k] > nm:
            flags[k] |= 1
        if k + 1 < n and i_next < len(troughs) and troughs[i_next] < peaks[k + 1]:
            if vm * (envelope[p] - nxt[k]) < (envelope[peaks[k + 1]] - nxt[k]):
                flags[k] |= 2
    return {"prev_amp": prev, "next_amp": nxt, "ratio": ratio, "flags": flags}


# --------------------------------------------------------------------------- a5
def calculate_bpm_series(peaks: np.ndarray, rate: int, params: Dict
                         ) -> Tuple[pd.Series, np.ndarray]:
    """bpm_analysis.py:1463-1484."""
    if len(peaks) < 2:
        return pd.Series(dtype=np.float64), np.array([])
    t = peaks / rate
    dt = np.diff(t)
    ok = dt > 1e-6
    if not np.any(ok):
        return pd.Series(dtype=np.float64), np.array([])
    inst = 60.0 / dt[ok]
    epoch = datetime.datetime.fromtimestamp(0)
    stamps = [epoch + datetime.timedelta(seconds=v) for v in t[1:][ok]]
    series = pd.Series(inst, index=stamps)
    if np.median(inst) > 0:
        win = f"{params['output_smoothing_window_sec']}s"
        smoothed = series.rolling(window=win, min_periods=1, center=True).mean()
    else:
        smoothed = pd.Series(dtype=np.float64)
    return smoothed, t[1:][ok]


# --------------------------------------------------------------------------- a6
def _steepest(series: pd.Series, window_sec: float, sign: int) -> Optional[Dict]:
    times = (series.index - series.index[0]).total_seconds()
    if times[-1] < window_sec:
        return None
    vals = series.values
    best_slope, best = 0, None
    for i in range(len(times) - 1):
        cand = np.where(times >= times[i] + window_sec)[0]
        if len(cand) == 0:
            break
        j = cand[0]
        dur = times[j] - times[i]
        if dur > 0:
            slope = (vals[j] - vals[i]) / dur
            if (slope < best_slope) if sign < 0 else (slope > best_slope):
                best_slope = slope
                best = {"start_time": series.index[i], "end_time": series.index[j],
                        "start_bpm": vals[i], "end_bpm": vals[j],
                        "slope_bpm_per_sec": slope, "duration_sec": dur}
    return best


def find_peak_recovery_rate(series: pd.Series, window_sec: int = 20) -> Optional[Dict]:
    """bpm_analysis.py:1552-1574 (searches from the BPM maximum onward)."""
    if series.empty or len(series) < 2:
        return None
    tail = series[series.idxmax():]
    if tail.empty:
        return None
    return _steepest(tail, window_sec, -1)


def find_peak_exertion_rate(series: pd.Series, window_sec: int = 20) -> Optional[Dict]:
    """bpm_analysis.py:1576-1595."""
    if series.empty or len(series) < 2:
        return None
    return _steepest(series, window_sec, +1)


# --------------------------------------------------------------------------- a7
def _hr_extrema(series: pd.Series, min_duration_sec: float):
    gaps = series.index.to_series().diff().dt.total_seconds()
    mean_gap = np.nanmean(gaps)
    dist = 5 if np.isnan(mean_gap) or mean_gap == 0 else int((min_duration_sec / 2) / mean_gap)
    tops, _ = find_peaks(series.values, prominence=5, distance=dist)
    bottoms, _ = find_peaks(-series.values, prominence=5, distance=dist)
    return tops, bottoms


def find_major_hr_inclines(series: pd.Series, min_duration_sec: int = 10,
                           min_bpm_increase: int = 15) -> List[Dict]:
    """bpm_analysis.py:1486-1516."""
    if series.empty or len(series) < 2:
        return []
    tops, bottoms = _hr_extrema(series, min_duration_sec)
    if len(tops) == 0 or len(bottoms) == 0:
        return []
    out = []
    v, ix = series.values, series.index
    for b in bottoms:
        after = tops[tops > b]
        if len(after) == 0:
            continue
        p = after[0]
        dur = (ix[p] - ix[b]).total_seconds()
        rise = v[p] - v[b]
        if dur >= min_duration_sec and rise >= min_bpm_increase:
            out.append({"start_time": ix[b], "end_time": ix[p], "start_bpm": v[b], "end_bpm": v[p],
                        "duration_sec": dur, "bpm_increase": rise, "slope_bpm_per_sec": rise / dur})
    out.sort(key=lambda d: d["slope_bpm_per_sec"], reverse=True)
    return out


def find_major_hr_declines(series: pd.Series, min_duration_sec: int = 10,
                           min_bpm_decrease: int = 15) -> List[Dict]:
    """bpm_analysis.py:1518-1550."""
    if series.empty or len(series) < 2:
        return []
    tops, bottoms = _hr_extrema(series, min_duration_sec)
    if len(tops) == 0 or len(bottoms) == 0:
        return []
    out = []
    v, ix = series.values, series.index
    for p in tops:
        after = bottoms[bottoms > p]
        if len(after) == 0:
            continue
        b = after[0]
        dur = (ix[b] - ix[p]).total_seconds()
        drop = v[p] - v[b]
        if dur >= min_duration_sec and drop >= min_bpm_decrease:
            out.append({"start_time": ix[p], "end_time": ix[b], "start_bpm": v[p], "end_bpm": v[b],
                        "duration_sec": dur, "bpm_decrease": drop,
                        "slope_bpm_per_sec": (v[b] - v[p]) / dur})
    out.sort(key=lambda d: d["slope_bpm_per_sec"])
    return out


# --------------------------------------------------------------------------- a8
def calculate_windowed_hrv(s1_peaks: np.ndarray, rate: int, params: Dict) -> pd.DataFrame:
    """bpm_analysis.py:1414-1461."""
    win = params["hrv_window_size_beats"]
    step = params["hrv_step_size_beats"]
    cols = ["time", "rmssdc", "sdnn", "bpm"]
    if len(s1_peaks) < win:
        return pd.DataFrame(columns=cols)
    rr = np.diff(s1_peaks) / rate
    t = s1_peaks / rate
    rows = []
    for i in range(0, len(rr) - win + 1, step):
        ms = rr[i:i + win] * 1000
        mid = (t[i] + t[i + win]) / 2.0
        mean_ms = np.mean(ms)
        sdnn = np.std(ms)
        rmssd = np.sqrt(np.mean(np.diff(ms) ** 2))
        mean_s = mean_ms / 1000.0
        rows.append({"time": mid, "rmssdc": rmssd / mean_s if mean_s > 0 else 0,
                     "sdnn": sdnn, "bpm": 60 / mean_s if mean_s > 0 else 0})
    if not rows:
        return pd.DataFrame(columns=cols)
    return pd.DataFrame(rows)


# ------------------------------------------------------------------ hrr / recovery phase
def calculate_hrr(smoothed_bpm_series: pd.Series, interval_sec: int = 60) -> Optional[Dict]:
    """bpm_analysis.py:1597-1610.  The abscissa of the interpolation is the index cast to int64 and
    floor-divided by 1e9 (:1606) -- seconds only for a datetime64[ns] index; the pandas installed
    here stores it as datetime64[us], and the restatement keeps that expression as it is."""
    if smoothed_bpm_series.empty or len(smoothed_bpm_series) < 2:
        return None
    peak_bpm, peak_time = smoothed_bpm_series.max(), smoothed_bpm_series.idxmax()                  # :1600
    check = peak_time + pd.Timedelta(seconds=interval_sec)                                         # :1601
    if check > smoothed_bpm_series.index.max():                                                    # :1602
        return None
    recovery_bpm = np.interp(check.timestamp(),                                                    # :1604-1607
                             (smoothed_bpm_series.index.astype(np.int64) // 10**9).to_numpy(dtype=float),
                             np.asarray(smoothed_bpm_series.values, dtype=float))
    return {"peak_bpm": peak_bpm, "peak_time": peak_time, "recovery_bpm": recovery_bpm,
            "recovery_check_time": check, "hrr_value_bpm": peak_bpm - recovery_bpm, "interval_sec": interval_sec}


def find_recovery_phase(bpm_series: pd.Series, bpm_times_sec: np.ndarray, params: Dict):
    """bpm_analysis.py:1612-1620."""
    if bpm_times_sec is None or len(bpm_times_sec) < 2:                                            # :1614
        return None, None
    peak_time_sec = bpm_times_sec[np.argmax(bpm_series.to_numpy())]                                # :1617
    return peak_time_sec, peak_time_sec + params.get("recovery_phase_duration_sec", 120.0)         # :1618


# ------------------------------------------------------------------ whole path
def front_end(audio_data: np.ndarray, sample_rate: int, params: Dict) -> Dict[str, object]:
    """a1..a4 chained the way analyze_wav_file does (:1731-1732, :1635)."""
    env, rate, filt = preprocess_pcm(audio_data, sample_rate, params)
    floor, troughs = calculate_dynamic_noise_floor(env, rate, params)
    peaks = find_raw_peaks(env, rate, params, floor.values)
    out = {"envelope": env, "rate": rate, "filtered": filt, "floor": floor.values,
           "troughs": np.asarray(troughs), "peaks": peaks}
    if len(peaks) >= 2:
        out.update(peak_metrics(env, rate, params, floor, peaks))
    return out


def beat_reductions(beats: np.ndarray, rate: int, params: Dict) -> Dict[str, object]:
    """a5..a8 chained the way _calculate_final_metrics does (:1701-1722)."""
    series, times = calculate_bpm_series(beats, rate, params)
    return {"smoothed_bpm": series, "bpm_times": times,
            "major_inclines": find_major_hr_inclines(series),
            "major_declines": find_major_hr_declines(series),
            "peak_recovery_stats": find_peak_recovery_rate(series),
            "peak_exertion_stats": find_peak_exertion_rate(series),
            "hrr_stats": calculate_hrr(series),
            "recovery_phase": find_recovery_phase(series, times, params),
            "windowed_hrv_df": calculate_windowed_hrv(beats, rate, params)}
