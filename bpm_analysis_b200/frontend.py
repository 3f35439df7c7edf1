"""Drop-in mirror of the data-parallel front end of the reference's ``bpm_analysis.py``.

Same names, signatures, return types and error behaviour as the reference
functions (cited per function), so ``analyze_wav_file``, ``gui.py`` and the
hugging-face-space app keep working when ``install()`` rebinds them.  Every
numeric step runs in libbpm_b200.so on the GPU; this module only adapts host
objects (WAV files, numpy arrays, pandas Series / DataFrame) at the boundary.
There is no CPU fallback: without the library or a CUDA device these raise.
"""
from __future__ import annotations

import datetime
import logging
import os
import warnings
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
from scipy.io import wavfile

from . import runtime
from .dropin import dropin
from .params import (HR_MIN_CHANGE_BPM, HR_MIN_DURATION_SEC, HR_PROMINENCE, SLOPE_WINDOW_SEC, band_edges,
                     effective_decimation)

__all__ = ["preprocess_audio", "preprocess_pcm", "read_wav", "_calculate_dynamic_noise_floor", "_find_raw_peaks",
           "_initialize_state", "calculate_bpm_series", "find_peak_recovery_rate", "find_peak_exertion_rate",
           "find_major_hr_inclines", "find_major_hr_declines", "calculate_windowed_hrv", "calculate_hrr",
           "find_recovery_phase", "find_peaks", "peak_trough_noise", "install"]


# --------------------------------------------------------------------------- a1
def preprocess_pcm(audio_data: np.ndarray, sample_rate: int, params: Dict, want_debug: bool = False,
                   want_filtered: bool = True):
    """Array-level body of ``preprocess_audio``: (envelope, rate, filtered | None, debug_int16 | None).

    The call also runs the noise-floor, raw-peak and per-peak-metric stages on the device
    (one ``bpm_stage_a``) and parks their results in a session keyed by the returned envelope
    array, so the calls ``analyze_wav_file`` makes next are served without another round trip
    (``dropin.py``).  ``want_filtered=False`` skips the band-passed signal (only tests and the
    debug WAV read it)."""
    ds, rate, clamped = effective_decimation(sample_rate, params)
    if clamped:                                                     # bpm_analysis.py:1023-1029
        _, highcut = band_edges(params)
        logging.warning(f"Original 'downsample_factor' of {params['downsample_factor']} is too high for a "
                        f"{highcut:g}Hz filter with a {sample_rate}Hz sample rate.")
        logging.warning(f"Adjusting 'downsample_factor' to a safe value of {ds}.")
    from .wav24 import S24Recording
    if not isinstance(audio_data, S24Recording):
        audio_data = np.asarray(audio_data)
    return dropin().preprocess(audio_data, int(sample_rate), params, bool(want_debug), bool(want_filtered))


def read_wav(file_path: str):
    """``wavfile.read`` (bpm_analysis.py:1014) without copying the file: the samples are memory-mapped
    and the decimate-first ingest (``bpm_host_gather_frames``) then touches one frame in ``ds``
    straight out of the page cache -- scipy's default read copies all of a 60-minute recording
    (345 MB, ~150 ms) to hand 2 MB of it to the filter.  Formats scipy cannot map (24-bit PCM) fall
    to ``wav24.map_s24`` -- the file mapped as bytes, the kept frames expanded to scipy's int32 by
    ``bpm_host_gather_s24`` -- and only what that does not recognise to the copying read; sample rate,
    dtype, channel layout, values and errors are scipy's either way."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            return wavfile.read(file_path, mmap=True)
        except ValueError:
            from .wav24 import map_s24
            mapped = map_s24(file_path)
            return mapped if mapped is not None else wavfile.read(file_path)


def preprocess_audio(file_path: str, params: Dict, output_directory: str) -> Tuple[np.ndarray, int]:
    """Reads, filters, and prepares the audio envelope.  Mirrors bpm_analysis.py:1007-1062."""
    save_debug_file = params["save_filtered_wav"]
    _ = params["downsample_factor"]                                 # KeyError like the reference
    sample_rate, audio_data = read_wav(file_path)
    envelope, rate, _, debug = preprocess_pcm(audio_data, sample_rate, params, want_debug=bool(save_debug_file),
                                              want_filtered=False)
    if save_debug_file:
        # the reference writes the same int16 signal twice (:1047-1050 and :1056-1060)
        wavfile.write(f"{os.path.splitext(file_path)[0]}_filtered_debug.wav", rate, debug)
        base_name = os.path.basename(os.path.splitext(file_path)[0])
        wavfile.write(os.path.join(output_directory, f"{base_name}_filtered_debug.wav"), rate, debug)
    return envelope, rate


def _as_signal(a) -> np.ndarray:
    """A contiguous float array WITHOUT copying what is one already: the session layer recognises the
    arrays it handed out by their memory (float32 ones included, ``output_dtype="float32"``)."""
    if isinstance(a, np.ndarray) and a.dtype in (np.float64, np.float32) and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


# --------------------------------------------------------------------------- a2
_index_cache: Dict[int, pd.Index] = {}


def _arange_index(n: int) -> pd.Index:
    """pd.Index(np.arange(n)) -- what the reference builds per call (:1076, :1082); an Index is
    immutable, so one per length is shared between the Series handed out."""
    ix = _index_cache.get(n)
    if ix is None:
        if len(_index_cache) > 8:
            _index_cache.clear()
        ix = _index_cache[n] = pd.Index(np.arange(n))
    return ix


def _calculate_dynamic_noise_floor(audio_envelope: np.ndarray, sample_rate: int, params: Dict
                                   ) -> Tuple[pd.Series, np.ndarray]:
    """Dynamic noise floor from sanitised troughs.  Mirrors bpm_analysis.py:1064-1117."""
    env = _as_signal(audio_envelope)
    d = dropin()
    floor, troughs, n_all, mode = d.noise_floor(env, int(sample_rate), params)
    if mode == 2:                                                   # :1073-1077
        logging.warning("Not enough troughs found for sanitization. Using a static noise floor.")
    else:
        logging.info(f"Trough Sanitization: Kept {len(troughs)} of {n_all} initial troughs.")     # :1099
        if mode == 1:                                               # :1107-1110
            logging.warning("Not enough sanitized troughs remaining. Using non-sanitized floor as fallback.")
    series = pd.Series(floor, index=_arange_index(len(env)), copy=False)
    d.note_floor_alias(env, floor, series.values, series)
    if mode != 2 and len(troughs) == 0:
        troughs = np.array([])                                      # np.array([]) of an empty list is float64 (:1117)
    return series, troughs


# --------------------------------------------------------------------------- a3 / a4
def _raw_peaks(envelope: np.ndarray, height_threshold: np.ndarray, sample_rate: int, params: Dict,
               with_metrics: bool):
    env, height = _as_signal(envelope), _as_signal(height_threshold)
    if height.shape != env.shape:                                   # what scipy's find_peaks raises
        raise ValueError("array size of lower interval border must match x")
    return dropin().raw_peaks(env, height, int(sample_rate), params, with_metrics)


_sorted_cache: "OrderedDict[tuple, tuple]" = OrderedDict()


def _sorted_list(values) -> list:
    """``sorted(values)`` (bpm_analysis.py:107) -- a list of the array's scalars in ascending order.
    The trough array we hand out is already ascending, and both classifier constructions get the
    same array: the list is built once per array (straight iteration when it is sorted already) and
    every caller receives its own shallow copy."""
    if not isinstance(values, np.ndarray) or values.ndim != 1:
        return sorted(values)
    key = (values.__array_interface__["data"][0], values.size, values.dtype.str)
    hit = _sorted_cache.get(key)
    if hit is not None and hit[0]() is values:
        return list(hit[1])
    asc = values.size < 2 or bool(np.all(values[:-1] <= values[1:]))
    lst = list(values) if asc else sorted(values)
    try:
        import weakref
        _sorted_cache[key] = (weakref.ref(values), lst)
        while len(_sorted_cache) > 4:
            _sorted_cache.popitem(last=False)
    except TypeError:
        pass
    return list(lst)


def _find_raw_peaks(self, height_threshold: np.ndarray) -> np.ndarray:
    """Replacement for ``PeakClassifier._find_raw_peaks`` (bpm_analysis.py:223-229)."""
    peaks, _ = _raw_peaks(self.audio_envelope, height_threshold, self.sample_rate, self.params, False)
    logging.info(f"Found {len(peaks)} raw peaks using dynamic height threshold.")
    return peaks


def _initialize_state(self, start_bpm_hint, precomputed_noise_floor, precomputed_troughs) -> Dict:
    """Replacement for ``PeakClassifier._initialize_state`` (bpm_analysis.py:85-111)."""
    state = {"analysis_data": {}}
    state["dynamic_noise_floor"], state["trough_indices"] = precomputed_noise_floor, precomputed_troughs
    peaks, met = _raw_peaks(self.audio_envelope, state["dynamic_noise_floor"].values, self.sample_rate,
                            self.params, True)
    logging.info(f"Found {len(peaks)} raw peaks using dynamic height threshold.")
    state["all_peaks"] = peaks
    state["analysis_data"]["dynamic_noise_floor_series"] = state["dynamic_noise_floor"]
    state["analysis_data"]["trough_indices"] = state["trough_indices"]
    deviation_times = (peaks[:-1] + peaks[1:]) / 2 / self.sample_rate
    state["smoothed_dev_series"] = pd.Series(met["smoothed"].copy(), index=pd.Index(deviation_times, copy=False), copy=False)
    state["analysis_data"]["deviation_series"] = state["smoothed_dev_series"]
    state["long_term_bpm"] = float(start_bpm_hint) if start_bpm_hint else 80.0
    state["candidate_beats"] = []
    state["beat_debug_info"] = {}
    state["long_term_bpm_history"] = []
    state["sorted_troughs"] = _sorted_list(state["trough_indices"])
    state["consecutive_rr_rejections"] = 0
    state["loop_idx"] = 0
    return state


def peak_trough_noise(audio_envelope: np.ndarray, noise_floor, raw_peaks: np.ndarray, trough_indices: np.ndarray,
                      params: Dict) -> Dict[str, np.ndarray]:
    """Surrounding-trough noise per raw peak (optional extra output, parity unpinned).

    The reference documents this check (Documentation/Changelog.md:454, "BPM Detection logic
    explained.md":262, :276-278) and keeps its parameters (config.py:30-31) but no longer has a
    function for it.  Returns ``prev_amp``, ``next_amp`` (envelope at the sanitised trough before /
    after each peak), ``ratio`` = deeper trough / floor at the peak and ``flags`` (bit 0: ratio >
    ``trough_noise_multiplier``; bit 1: look-ahead veto with ``trough_veto_multiplier``)."""
    env = np.ascontiguousarray(audio_envelope, dtype=np.float64)
    floor = np.ascontiguousarray(getattr(noise_floor, "values", noise_floor), dtype=np.float64)
    if floor.shape != env.shape:
        raise ValueError("noise floor and envelope must have the same length")
    return runtime.ops().peak_trough_noise(env, floor, np.asarray(raw_peaks, dtype=np.int64),
                                           np.sort(np.asarray(trough_indices, dtype=np.int64)),
                                           float(params.get("trough_noise_multiplier", 3.0)),
                                           float(params.get("trough_veto_multiplier", 2.1)))


def find_peaks(x, height=None, prominence=None, distance=None):
    """scipy.signal.find_peaks-shaped operator (the subset the reference uses), on the GPU."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    if distance is not None and distance < 1:
        raise ValueError("`distance` must be greater or equal to 1")
    h = None
    if height is not None:
        h = np.broadcast_to(np.asarray(height, dtype=np.float64), x.shape).copy()
    d = 1 if distance is None else int(np.ceil(distance))
    if len(x) < 3:
        return np.array([], dtype=np.int64), {}
    return runtime.ops().find_peaks(x, h, None if prominence is None else float(prominence), d, 1), {}


# --------------------------------------------------------------------------- a5
def _epoch_us() -> int:
    """datetime.fromtimestamp(0) -- the reference's series epoch (local time, :1472)."""
    e = datetime.datetime.fromtimestamp(0)
    return int((e - datetime.datetime(1970, 1, 1)) // datetime.timedelta(microseconds=1))


def _series_key(series: pd.Series) -> tuple:
    v = np.ascontiguousarray(series.values, dtype=np.float64)
    return (len(v), hash(v.tobytes()), hash(series.index.asi8.tobytes()) if len(v) else 0)


_series_results: "OrderedDict[tuple, dict]" = OrderedDict()


def _beat_results_of(series: pd.Series) -> Optional[dict]:
    """The a5..a8 results computed together with ``series`` (``calculate_bpm_series`` registers
    every Series it hands out by content), or None for a Series from somewhere else."""
    if not isinstance(series.index, pd.DatetimeIndex):
        return None
    return _series_results.get(_series_key(series))


def calculate_bpm_series(peaks: np.ndarray, sample_rate: int, params: Dict) -> Tuple[pd.Series, np.ndarray]:
    """Smoothed BPM series from S1 peaks.  Mirrors bpm_analysis.py:1463-1484.

    The same device round trip computes the steepest slopes, the extrema of the smoothed series
    and the windowed HRV table of this beat list (``dropin.beat_metrics``); the reductions that
    ``_calculate_final_metrics`` applies next (:1704-1710) find them by the Series they are given."""
    if len(peaks) < 2:
        return pd.Series(dtype=np.float64), np.array([])
    res = dropin().beat_metrics(np.asarray(peaks, dtype=np.int64), int(sample_rate), params)
    if res.get("n_series", 0) == 0:
        return pd.Series(dtype=np.float64), np.array([])
    inst, smoothed, times, us = res["inst"], res["smoothed"], res["times"], res["stamps"]
    if not (np.median(inst) > 0):
        return pd.Series(dtype=np.float64), times.copy()
    index = pd.DatetimeIndex((us + _epoch_us()).astype("datetime64[us]"))
    series = pd.Series(smoothed.copy(), index=index, copy=False)
    _series_results[_series_key(series)] = res
    while len(_series_results) > 16:
        _series_results.popitem(last=False)
    return series, times.copy()


# --------------------------------------------------------------------------- a6
def _series_arrays(series: pd.Series):
    return np.ascontiguousarray(series.values, dtype=np.float64), series.index.as_unit("us").asi8


def _slope_record(series: pd.Series, vals, us, hit, origin: int) -> Dict:
    i, j, slope = hit
    dur = float((us[j] - us[origin]) / 1e6 - (us[i] - us[origin]) / 1e6)
    ix = series.index
    return {"start_time": ix[i], "end_time": ix[j], "start_bpm": vals[i], "end_bpm": vals[j],
            "slope_bpm_per_sec": slope, "duration_sec": dur}


def _steepest(series: pd.Series, sign: int, window_sec):
    vals, us = _series_arrays(series)
    res = _beat_results_of(series) if float(window_sec) == float(SLOPE_WINDOW_SEC) else None
    if res is not None:
        r = res["slopes"][0:4] if sign > 0 else res["slopes"][4:8]
        hit = None if r[0] == 0.0 else (int(r[1]), int(r[2]), float(r[3]))
    else:
        hit = runtime.ops().steepest(vals, us, sign, float(window_sec))
    return vals, us, hit


def find_peak_recovery_rate(smoothed_bpm_series: pd.Series, window_sec: int = SLOPE_WINDOW_SEC) -> Optional[Dict]:
    """Steepest HR decline after the peak BPM.  Mirrors bpm_analysis.py:1552-1574."""
    if smoothed_bpm_series.empty or len(smoothed_bpm_series) < 2:
        return None
    vals, us, hit = _steepest(smoothed_bpm_series, -1, window_sec)
    if hit is None:
        return None
    return _slope_record(smoothed_bpm_series, vals, us, hit, int(np.argmax(vals)))


def find_peak_exertion_rate(smoothed_bpm_series: pd.Series, window_sec: int = SLOPE_WINDOW_SEC) -> Optional[Dict]:
    """Steepest HR increase over the recording.  Mirrors bpm_analysis.py:1576-1595."""
    if smoothed_bpm_series.empty or len(smoothed_bpm_series) < 2:
        return None
    vals, us, hit = _steepest(smoothed_bpm_series, +1, window_sec)
    if hit is None:
        return None
    return _slope_record(smoothed_bpm_series, vals, us, hit, 0)


# --------------------------------------------------------------------------- a7
def _hr_extrema(series: pd.Series, min_duration_sec: float):
    vals, us = _series_arrays(series)
    gaps = np.diff(us) / 1e6                                        # index.to_series().diff().dt.total_seconds()
    # np.nanmean over [NaN, gaps...]: NaN -> 0, summed with it, divided by the non-NaN count
    mean_gap = np.sum(np.concatenate([[0.0], gaps])) / len(gaps) if len(gaps) else np.nan
    dist = 5 if np.isnan(mean_gap) or mean_gap == 0 else int((min_duration_sec / 2) / mean_gap)
    if dist < 1:
        raise ValueError("`distance` must be greater or equal to 1")
    res = _beat_results_of(series)
    if res is not None and res.get("hr_distance") == dist and "tops" in res:
        return vals, res["tops"], res["bottoms"]
    o = runtime.ops()
    tops = o.find_peaks(vals, None, float(HR_PROMINENCE), dist, +1)
    bottoms = o.find_peaks(vals, None, float(HR_PROMINENCE), dist, -1)
    return vals, tops, bottoms


def find_major_hr_inclines(smoothed_bpm_series: pd.Series, min_duration_sec: int = HR_MIN_DURATION_SEC,
                           min_bpm_increase: int = HR_MIN_CHANGE_BPM) -> List[Dict]:
    """Sustained HR increases.  Mirrors bpm_analysis.py:1486-1516."""
    if smoothed_bpm_series.empty or len(smoothed_bpm_series) < 2:
        return []
    vals, tops, bottoms = _hr_extrema(smoothed_bpm_series, min_duration_sec)
    if len(bottoms) == 0 or len(tops) == 0:
        return []
    ix = smoothed_bpm_series.index
    out = []
    for b in bottoms:
        k = np.searchsorted(tops, b, side="right")
        if k >= len(tops):
            continue
        p = tops[k]
        dur = (ix[p] - ix[b]).total_seconds()
        rise = vals[p] - vals[b]
        if dur >= min_duration_sec and rise >= min_bpm_increase:
            out.append({"start_time": ix[b], "end_time": ix[p], "start_bpm": vals[b], "end_bpm": vals[p],
                        "duration_sec": dur, "bpm_increase": rise, "slope_bpm_per_sec": rise / dur})
    out.sort(key=lambda d: d["slope_bpm_per_sec"], reverse=True)
    return out


def find_major_hr_declines(smoothed_bpm_series: pd.Series, min_duration_sec: int = HR_MIN_DURATION_SEC,
                           min_bpm_decrease: int = HR_MIN_CHANGE_BPM) -> List[Dict]:
    """Sustained HR decreases.  Mirrors bpm_analysis.py:1518-1550."""
    if smoothed_bpm_series.empty or len(smoothed_bpm_series) < 2:
        return []
    vals, tops, bottoms = _hr_extrema(smoothed_bpm_series, min_duration_sec)
    if len(bottoms) == 0 or len(tops) == 0:
        return []
    ix = smoothed_bpm_series.index
    out = []
    for p in tops:
        k = np.searchsorted(bottoms, p, side="right")
        if k >= len(bottoms):
            continue
        b = bottoms[k]
        dur = (ix[b] - ix[p]).total_seconds()
        drop = vals[p] - vals[b]
        if dur >= min_duration_sec and drop >= min_bpm_decrease:
            out.append({"start_time": ix[p], "end_time": ix[b], "start_bpm": vals[p], "end_bpm": vals[b],
                        "duration_sec": dur, "bpm_decrease": drop, "slope_bpm_per_sec": (vals[b] - vals[p]) / dur})
    out.sort(key=lambda d: d["slope_bpm_per_sec"])
    return out


# --------------------------------------------------------------------------- a8
def calculate_windowed_hrv(s1_peaks: np.ndarray, sample_rate: int, params: Dict) -> pd.DataFrame:
    """Sliding-window SDNN / RMSSDc.  Mirrors bpm_analysis.py:1414-1461."""
    window_size_beats = params["hrv_window_size_beats"]
    step_size_beats = params["hrv_step_size_beats"]
    cols = ["time", "rmssdc", "sdnn", "bpm"]
    if len(s1_peaks) < window_size_beats:
        logging.warning(f"Not enough beats ({len(s1_peaks)}) to perform windowed HRV analysis with a window of "
                        f"{window_size_beats} beats.")
        return pd.DataFrame(columns=cols)
    res = dropin().beat_metrics(np.asarray(s1_peaks, dtype=np.int64), int(sample_rate), params)
    rows = res.get("hrv", np.zeros((0, 4)))
    if len(rows) == 0:
        logging.warning("Could not perform windowed HRV analysis. Recording may be too short or have too few beats.")
        return pd.DataFrame(columns=cols)
    logging.info(f"Beat-based windowed HRV analysis complete. Generated {len(rows)} data points.")
    return pd.DataFrame(rows.copy(), columns=cols)


# --------------------------------------------------------------------------- hrr / recovery phase
def calculate_hrr(smoothed_bpm_series: pd.Series, interval_sec: int = 60) -> Optional[Dict]:
    """Heart-rate recovery over a fixed interval after the peak.  Mirrors bpm_analysis.py:1597-1610,
    including what it does with the index under the installed pandas: the interpolation abscissa is
    ``index.astype(int64) // 10**9`` -- seconds only while the index unit is ns (pandas 3 stores
    these stamps as datetime64[us]; SURVEY.md section 8c "version hazards").  Host glue on a
    series of one value per beat: an arg-max and one np.interp."""
    if smoothed_bpm_series.empty or len(smoothed_bpm_series) < 2:
        return None
    peak_time = smoothed_bpm_series.idxmax()
    peak_bpm = smoothed_bpm_series.max()
    check_time = peak_time + pd.Timedelta(seconds=interval_sec)
    if check_time > smoothed_bpm_series.index.max():
        return None
    xp = (smoothed_bpm_series.index.astype(np.int64) // 10**9).to_numpy(dtype=float)
    fp = np.asarray(smoothed_bpm_series.values, dtype=float)
    recovery_bpm = np.interp(check_time.timestamp(), xp, fp)
    return {"peak_bpm": peak_bpm, "peak_time": peak_time, "recovery_bpm": recovery_bpm,
            "recovery_check_time": check_time, "hrr_value_bpm": peak_bpm - recovery_bpm, "interval_sec": interval_sec}


def find_recovery_phase(bpm_series: pd.Series, bpm_times_sec: np.ndarray, params: Dict
                        ) -> Tuple[Optional[float], Optional[float]]:
    """Peak of the preliminary BPM series and the end of the high-contractility window after it.
    Mirrors bpm_analysis.py:1612-1620."""
    if bpm_times_sec is None or len(bpm_times_sec) < 2:
        logging.warning("Not enough preliminary beats to determine a recovery phase.")
        return None, None
    peak_time_sec = bpm_times_sec[np.argmax(bpm_series.to_numpy())]
    recovery_end_time_sec = peak_time_sec + params.get("recovery_phase_duration_sec", 120.0)
    logging.info(f"Peak BPM detected in preliminary pass at {peak_time_sec:.2f}s. High-contractility state defined "
                 f"until {recovery_end_time_sec:.2f}s.")
    return peak_time_sec, recovery_end_time_sec


# --------------------------------------------------------------------------- install
def install(ref_module, classifier: bool = True, corrections: bool = True, reports: bool = True):
    """Rebind the reference module's front-end names to this package.

    ``ref_module`` is an imported reference ``bpm_analysis`` module.  Callers that did
    ``from bpm_analysis import analyze_wav_file`` keep working because that function
    resolves these names from the module's globals at call time.  With ``classifier=True``
    (default) ``PeakClassifier.classify_peaks`` is rebound to the compiled sequential loop
    (``classifier.py`` -> libbpm_host.so) as well; with ``corrections=True`` (default) so are
    ``correct_peaks_by_rhythm`` and ``_fix_rhythmic_discontinuities`` (``corrections.py``); with
    ``reports=True`` (default) ``ReportGenerator`` is the index-based writer of ``reports.py`` (same files).
    """
    for name in ("preprocess_audio", "_calculate_dynamic_noise_floor", "calculate_bpm_series",
                 "find_peak_recovery_rate", "find_peak_exertion_rate", "find_major_hr_inclines",
                 "find_major_hr_declines", "calculate_windowed_hrv", "calculate_hrr", "find_recovery_phase"):
        setattr(ref_module, name, globals()[name])
    ref_module.PeakClassifier._find_raw_peaks = _find_raw_peaks
    ref_module.PeakClassifier._initialize_state = _initialize_state
    if classifier:
        from . import classifier as _classifier
        _classifier.install(ref_module)
    if corrections:
        from . import corrections as _corrections
        _corrections.install(ref_module)
    if reports:
        from . import reports as _reports
        _reports.install_reports(ref_module)
    return ref_module
