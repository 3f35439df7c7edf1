"""Drop-in replacement for ``PeakClassifier.classify_peaks`` (bpm_analysis.py:113-131) backed by
compiled host code (``csrc/classifier.cpp`` -> ``libbpm_host.so``, ``include/bpm_host.h``).

SURVEY.md section 8f, rank 1: once the numeric front end runs on the GPU the reference's
sequential S1/S2 loop (run twice per file, ~65 us of Python per raw peak) is what is left.  The
loop carries state from every decision to the next, so it is host code, not a kernel; this
module only marshals the state the reference's constructor prepared (``_initialize_state``,
served by ``frontend``) and turns the result back into the reference's Python objects: the same
``final_peaks``, the same ``beat_debug_info`` strings, the same ``long_term_bpm_series``.

There is no Python fallback: without ``libbpm_host.so`` this raises.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import threading
from typing import Dict, Tuple

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.environ.get("BPM_HOST_LIB") or os.path.join(HERE, "libbpm_host.so")
HOST_ABI_VERSION = 2

PEAK_TYPE_LABELS = {1: "S1 (Paired)", 2: "S2 (Paired)", 3: "Lone S1", 4: "Lone S1 (Corrected by Cascade Reset)",
                    5: "Lone S1 (Last Peak)", 6: "Noise"}

# (field, params key, reference default or None = required -> KeyError like the reference)
_DOUBLE_FIELDS = (
    ("pairing_confidence_threshold", "pairing_confidence_threshold", None),
    ("contractility_bpm_low", "contractility_bpm_low", None),
    ("contractility_bpm_high", "contractility_bpm_high", None),
    ("s1_s2_interval_cap_sec", "s1_s2_interval_cap_sec", None),
    ("s1_s2_interval_rr_fraction", "s1_s2_interval_rr_fraction", None),
    ("interval_penalty_start_factor", "interval_penalty_start_factor", 1.0),
    ("interval_penalty_full_factor", "interval_penalty_full_factor", 1.4),
    ("interval_max_penalty", "interval_max_penalty", 0.75),
    ("kickstart_check_threshold", "kickstart_check_threshold", 0.3),
    ("stability_confidence_floor", "stability_confidence_floor", 0.85),
    ("stability_confidence_ceiling", "stability_confidence_ceiling", 1.10),
    ("s2_s1_ratio_low_bpm", "s2_s1_ratio_low_bpm", None),
    ("s2_s1_ratio_high_bpm", "s2_s1_ratio_high_bpm", None),
    ("penalty_amount_min", "penalty_amount_min", 0.15),
    ("penalty_amount_max", "penalty_amount_max", 0.40),
    ("s1_s2_boost_ratio", "s1_s2_boost_ratio", 1.2),
    ("boost_amount_min", "boost_amount_min", 0.10),
    ("boost_amount_max", "boost_amount_max", 0.35),
    ("lone_s1_confidence_threshold", "lone_s1_confidence_threshold", 0.6),
    ("lone_s1_forward_check_pct", "lone_s1_forward_check_pct", 0.6),
    ("lone_s1_rhythm_weight", "lone_s1_rhythm_weight", 0.65),
    ("lone_s1_amplitude_weight", "lone_s1_amplitude_weight", 0.35),
    ("cascade_reset_trigger_count", "cascade_reset_trigger_count", 3),
    ("min_bpm", "min_bpm", None),
    ("max_bpm", "max_bpm", None),
)


class ClassifierParams(C.Structure):
    _fields_ = ([(f, C.c_double) for f, _, _ in _DOUBLE_FIELDS] +
                [("start_bpm", C.c_double), ("peak_bpm_time_sec", C.c_double), ("recovery_end_time_sec", C.c_double),
                 ("has_recovery_window", C.c_int32), ("enable_interval_penalty", C.c_int32),
                 ("stability_history_window", C.c_int32), ("flags", C.c_int32)])


CLASSIFY_NO_TEXT = 1


class ClassifyJob(C.Structure):
    _fields_ = [("envelope", C.c_void_p), ("noise_floor", C.c_void_p), ("m", C.c_int64), ("raw_peaks", C.c_void_p),
                ("n_peaks", C.c_int64), ("dev_times", C.c_void_p), ("dev_values", C.c_void_p), ("n_dev", C.c_int64),
                ("sample_rate", C.c_double), ("params", C.POINTER(ClassifierParams))]


class ClassifierEvent(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("pad", C.c_int32)]


class Classification(C.Structure):
    _fields_ = [("n_peaks", C.c_int64), ("n_beats", C.c_int64), ("n_history", C.c_int64), ("n_events", C.c_int64),
                ("text_bytes", C.c_int64), ("beat_positions", C.POINTER(C.c_int64)),
                ("peak_types", C.POINTER(C.c_int32)), ("text_offsets", C.POINTER(C.c_int64)),
                ("text", C.POINTER(C.c_char)), ("history_times", C.POINTER(C.c_double)),
                ("history_bpm", C.POINTER(C.c_double)), ("events", C.POINTER(ClassifierEvent)),
                ("final_long_term_bpm", C.c_double), ("final_consecutive_rr_rejections", C.c_int64)]


EXPORTED_SYMBOLS = ("bpm_host_abi_version", "bpm_host_format_fixed", "bpm_classify_peaks", "bpm_classify_peaks_batch",
                    "bpm_classification_free", "bpm_host_threads", "bpm_host_gather_frames", "bpm_host_gather_s24",
                    "bpm_host_format_rows")


class HostLibraryError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def load_host_library(path: str = HOST_LIB_PATH):
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(path):
            raise HostLibraryError(f"{path} not found: build it with `python -m bpm_analysis_b200.build` "
                                   "(there is no Python fallback for the classifier)")
        lib = C.CDLL(path)
        for name in EXPORTED_SYMBOLS:
            if not hasattr(lib, name):
                raise HostLibraryError(f"{path} does not export {name}")
        lib.bpm_host_abi_version.restype, lib.bpm_host_abi_version.argtypes = C.c_int, []
        lib.bpm_classify_peaks.restype = C.c_int
        lib.bpm_classify_peaks.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_double, C.POINTER(ClassifierParams),
                                           C.POINTER(C.POINTER(Classification))]
        lib.bpm_classification_free.restype, lib.bpm_classification_free.argtypes = None, [C.POINTER(Classification)]
        lib.bpm_classify_peaks_batch.restype = C.c_int
        lib.bpm_classify_peaks_batch.argtypes = [C.POINTER(ClassifyJob), C.c_int64, C.c_int,
                                                 C.POINTER(C.POINTER(Classification)), C.POINTER(C.c_int)]
        lib.bpm_host_threads.restype, lib.bpm_host_threads.argtypes = C.c_int, []
        lib.bpm_host_gather_frames.restype = C.c_int
        lib.bpm_host_gather_frames.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        lib.bpm_host_gather_s24.restype = C.c_int
        lib.bpm_host_gather_s24.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        lib.bpm_host_format_rows.restype = C.c_int64
        lib.bpm_host_format_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_char_p, C.c_char_p,
                                             C.c_char_p, C.c_int, C.c_void_p, C.c_int64]
        lib.bpm_host_format_fixed.restype = C.c_int
        lib.bpm_host_format_fixed.argtypes = [C.c_double, C.c_int, C.c_char_p, C.c_size_t]
        if lib.bpm_host_abi_version() != HOST_ABI_VERSION:
            raise HostLibraryError(f"ABI mismatch: library {lib.bpm_host_abi_version()}, binding {HOST_ABI_VERSION}")
        _lib = lib
        return lib


def pack_params(params: Dict, start_bpm: float, peak_bpm_time_sec, recovery_end_time_sec,
                with_text: bool = True) -> ClassifierParams:
    """DEFAULT_PARAMS (config.py) -> BpmClassifierParams, with the reference's ``.get`` defaults.
    ``with_text=False`` skips the per-peak debug strings (decisions are unchanged)."""
    p = ClassifierParams()
    for field, key, default in _DOUBLE_FIELDS:
        setattr(p, field, float(params[key] if default is None else params.get(key, default)))
    p.start_bpm = float(start_bpm)
    p.has_recovery_window = int(peak_bpm_time_sec is not None and recovery_end_time_sec is not None)
    p.peak_bpm_time_sec = float(peak_bpm_time_sec) if p.has_recovery_window else 0.0
    p.recovery_end_time_sec = float(recovery_end_time_sec) if p.has_recovery_window else 0.0
    p.enable_interval_penalty = int(bool(params.get("enable_interval_penalty", True)))
    p.stability_history_window = int(params.get("stability_history_window", 20))
    p.flags = 0 if with_text else CLASSIFY_NO_TEXT
    return p


def _as(ptr, n: int, dtype) -> np.ndarray:
    if n == 0:
        return np.empty(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def classify_arrays(envelope: np.ndarray, noise_floor: np.ndarray, raw_peaks: np.ndarray, dev_times: np.ndarray,
                    dev_values: np.ndarray, sample_rate, packed: ClassifierParams) -> Dict[str, object]:
    """Array-level call of ``bpm_classify_peaks``: returns beat positions, per-peak types and debug
    strings, the long-term BPM trace and the events the reference logs."""
    lib = load_host_library()
    env = np.ascontiguousarray(envelope, dtype=np.float64)
    floor = np.ascontiguousarray(noise_floor, dtype=np.float64)
    peaks = np.ascontiguousarray(raw_peaks, dtype=np.int64)
    dt = np.ascontiguousarray(dev_times, dtype=np.float64)
    dv = np.ascontiguousarray(dev_values, dtype=np.float64)
    if floor.shape != env.shape or dt.shape != dv.shape:
        raise ValueError("envelope / noise floor and deviation index / values must have matching lengths")
    out = C.POINTER(Classification)()
    rc = lib.bpm_classify_peaks(env.ctypes.data, floor.ctypes.data, env.shape[0], peaks.ctypes.data, peaks.shape[0],
                                dt.ctypes.data, dv.ctypes.data, dt.shape[0], float(sample_rate), C.byref(packed),
                                C.byref(out))
    if rc != 0:
        raise ValueError(f"bpm_classify_peaks rejected its arguments (code {rc}): raw peaks must be >= 2, strictly "
                         "ascending and inside the envelope; stability_history_window >= 1")
    try:
        return _unpack(out.contents)
    finally:
        lib.bpm_classification_free(out)


def _unpack(r: Classification) -> Dict[str, object]:
    n = int(r.n_peaks)
    texts = C.string_at(r.text, int(r.text_bytes)).decode("utf-8").split("\x00")[:n]   # NUL-terminated entries
    return {
        "beat_positions": _as(r.beat_positions, int(r.n_beats), np.int64),
        "peak_types": _as(r.peak_types, n, np.int32),
        "texts": texts,
        "history_times": _as(r.history_times, int(r.n_history), np.float64),
        "history_bpm": _as(r.history_bpm, int(r.n_history), np.float64),
        "events": [(int(r.events[i].kind), int(r.events[i].a), int(r.events[i].b)) for i in range(int(r.n_events))],
        "final_long_term_bpm": float(r.final_long_term_bpm),
        "final_consecutive_rr_rejections": int(r.final_consecutive_rr_rejections),
    }


def classify_batch(jobs, n_threads: int = 0):
    """Classify independent recordings on host threads (``bpm_classify_peaks_batch``).

    ``jobs``: sequence of ``(envelope, noise_floor, raw_peaks, dev_times, dev_values, sample_rate,
    packed_params)``.  Returns one result dict per job (as ``classify_arrays``); raises
    ``ValueError`` naming the first job the library rejected."""
    lib = load_host_library()
    n = len(jobs)
    arr = (ClassifyJob * n)()
    keep = []
    for i, (env, floor, peaks, dt, dv, rate, packed) in enumerate(jobs):
        a = [np.ascontiguousarray(env, dtype=np.float64), np.ascontiguousarray(floor, dtype=np.float64),
             np.ascontiguousarray(peaks, dtype=np.int64), np.ascontiguousarray(dt, dtype=np.float64),
             np.ascontiguousarray(dv, dtype=np.float64)]
        if a[1].shape != a[0].shape or a[3].shape != a[4].shape:
            raise ValueError(f"job {i}: envelope / noise floor and deviation index / values must have matching lengths")
        keep.append((a, packed))
        arr[i] = ClassifyJob(a[0].ctypes.data, a[1].ctypes.data, a[0].shape[0], a[2].ctypes.data, a[2].shape[0],
                             a[3].ctypes.data, a[4].ctypes.data, a[3].shape[0], float(rate), C.pointer(packed))
    outs = (C.POINTER(Classification) * n)()
    status = (C.c_int * n)()
    rc = lib.bpm_classify_peaks_batch(arr, n, int(n_threads), outs, status)
    try:
        if rc != 0:
            raise ValueError(f"bpm_classify_peaks_batch failed (code {rc})")
        for i in range(n):
            if status[i] != 0:
                raise ValueError(f"job {i}: bpm_classify_peaks rejected its arguments (code {status[i]})")
        return [_unpack(outs[i].contents) for i in range(n)]
    finally:
        for i in range(n):
            if outs[i]:
                lib.bpm_classification_free(outs[i])


def classify_peaks(self) -> Tuple[np.ndarray, np.ndarray, Dict]:
    """Replacement for ``PeakClassifier.classify_peaks`` (bpm_analysis.py:113-131): same return
    value, same ``self.state`` afterwards."""
    st = self.state
    all_peaks = st["all_peaks"]
    if len(all_peaks) < 2:                                          # :115-116
        return all_peaks, all_peaks, {"beat_debug_info": {}}
    if st["loop_idx"] == 0 and not st["candidate_beats"]:
        # Python-level errors the reference's first iteration raises before any decision is made
        # (:135-141 with an empty window, :1136 with equal contractility anchors and a Python-float BPM)
        if self.params.get("stability_history_window", 20) == 0:
            raise ZeroDivisionError("division by zero")
        _ = ((st["long_term_bpm"] - self.params["contractility_bpm_low"]) /
             (self.params["contractility_bpm_high"] - self.params["contractility_bpm_low"]))
        packed = pack_params(self.params, st["long_term_bpm"], self.peak_bpm_time_sec, self.recovery_end_time_sec)
        dev = st["smoothed_dev_series"]
        res = classify_arrays(self.audio_envelope, st["dynamic_noise_floor"].values, all_peaks,
                              dev.index.values, dev.values, self.sample_rate, packed)
        keys = list(all_peaks)                                      # np.int64 scalars, like the reference's keys
        for kind, a, b in res["events"]:                            # the reference logs these as it goes
            if kind == 1:
                override = self.params.get("kickstart_override_ratio", 0.6)
                logging.info(f"KICK-START: Found {a}/{b} S1->Noise patterns. Overriding pairing ratio to {override}.")
                st["pairing_ratio_override"] = override
            elif kind == 2:
                logging.info(f"CASCADE RESET: Forcing peak at {keys[a] / self.sample_rate:.2f}s as Lone S1 due to "
                             f"repeated rhythmic failures.")
        st["candidate_beats"] = [keys[i] for i in res["beat_positions"]]
        st["beat_debug_info"] = dict(zip(keys, res["texts"]))
        st["long_term_bpm_history"] = list(zip(res["history_times"], res["history_bpm"]))
        st["long_term_bpm"] = res["final_long_term_bpm"]
        st["consecutive_rr_rejections"] = res["final_consecutive_rr_rejections"]
        st["loop_idx"] = len(all_peaks)
    # _finalize_results, :214-221
    final_peaks = np.array(sorted(list(dict.fromkeys(st["candidate_beats"]))))
    st["analysis_data"]["beat_debug_info"] = st["beat_debug_info"]
    if st["long_term_bpm_history"]:
        lt_times, lt_values = zip(*st["long_term_bpm_history"])
        st["analysis_data"]["long_term_bpm_series"] = pd.Series(lt_values, index=lt_times)
    return final_peaks, st["all_peaks"], st["analysis_data"]


def install(ref_module):
    """Rebind ``PeakClassifier.classify_peaks`` of an imported reference module."""
    ref_module.PeakClassifier.classify_peaks = classify_peaks
    return ref_module
