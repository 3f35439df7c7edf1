"""The params contract of the hot path.

The reference passes one flat dict (``config.py:3-108`` ``DEFAULT_PARAMS``) to
every function.  This package consumes that dict unchanged; the table below
only lists the keys the data-parallel front end reads (12 + 2 for the optional trough-noise metric) with
the reference's default values, so the package is usable standalone.  Callers
that already hold the reference's ``DEFAULT_PARAMS`` just pass it through.

Optional keys this package adds (absent => reference behaviour):

``filter_mode``   ``"parity"`` (default): stride-decimate, then band-pass at the
                  envelope rate -- what ``bpm_analysis.py:1031-1045`` does.
                  ``"fullrate"``: band-pass at the original rate (the order
                  ``README.md:6`` documents), then take every ds-th sample.
``output_dtype``  ``"float64"`` (default) or ``"float32"``: the "float32 mode" of the north star.
                  All stages compute in float64; the envelope / noise floor / per-peak series
                  handed back to the host are rounded to float32 (half the read-back bytes;
                  trough / peak / beat lists are unchanged).
``lowcut_hz`` / ``highcut_hz``   band edges, hard-coded 20/150 in
                  ``bpm_analysis.py:1018``; exposed for the C5 sweep.
"""
from __future__ import annotations

from typing import Dict

HOT_PATH_DEFAULTS: Dict[str, object] = {
    "downsample_factor": 300,
    "save_filtered_wav": True,
    "min_peak_distance_sec": 0.05,
    "peak_prominence_quantile": 0.1,
    "trough_prominence_quantile": 0.1,
    "noise_floor_quantile": 0.20,
    "noise_window_sec": 10,
    "trough_rejection_multiplier": 4.0,
    "deviation_smoothing_factor": 0.05,
    "output_smoothing_window_sec": 5,
    "hrv_window_size_beats": 40,
    "hrv_step_size_beats": 5,
    # read only by the optional surrounding-trough noise metric (frontend.peak_trough_noise)
    "trough_veto_multiplier": 2.1,
    "trough_noise_multiplier": 3.0,
}

# keys the sequential classifier reads (classifier.py -> libbpm_host.so), at the values of the
# reference's DEFAULT_PARAMS (config.py); keys absent there fall back to the .get() defaults in
# bpm_analysis.py (cascade_reset_trigger_count 3, enable_interval_penalty True)
CLASSIFIER_DEFAULTS: Dict[str, object] = {
    "pairing_confidence_threshold": 0.5,
    "contractility_bpm_low": 120.0,
    "contractility_bpm_high": 140.0,
    "s1_s2_interval_cap_sec": 0.4,
    "s1_s2_interval_rr_fraction": 0.7,
    "interval_penalty_start_factor": 1.0,
    "interval_penalty_full_factor": 1.4,
    "interval_max_penalty": 0.75,
    "kickstart_check_threshold": 0.3,
    "kickstart_override_ratio": 0.6,
    "stability_history_window": 20,
    "stability_confidence_floor": 0.6,
    "stability_confidence_ceiling": 1.25,
    "s2_s1_ratio_low_bpm": 1.5,
    "s2_s1_ratio_high_bpm": 1.1,
    "penalty_amount_min": 0.1,
    "penalty_amount_max": 0.3,
    "s1_s2_boost_ratio": 1.2,
    "boost_amount_min": 0.1,
    "boost_amount_max": 0.35,
    "lone_s1_confidence_threshold": 0.5,
    "lone_s1_forward_check_pct": 0.5,
    "lone_s1_rhythm_weight": 0.65,
    "lone_s1_amplitude_weight": 0.35,
    "min_bpm": 40,
    "max_bpm": 240,
    # correction passes (corrections.py)
    "rr_correction_threshold_pct": 0.4,
    "rr_correction_long_interval_pct": 1.7,
    "penalty_waiver_strength_ratio": 4.0,
    "penalty_waiver_max_s2_s1_ratio": 2.5,
}

# constants the reference hard-codes on the hot path (file:line in bpm_analysis.py)
LOWCUT_HZ = 20.0            # :1018
HIGHCUT_HZ = 150.0          # :1018
FILTER_ORDER = 2            # :1044
MIN_TROUGHS_FOR_DYNAMIC = 5  # :1073
ROLLING_MIN_PERIODS = 3     # :1085
MIN_SANITIZED_TROUGHS = 2   # :1102  (strictly more than)
FALLBACK_QUANTILE = 0.1     # :1114
STRENGTH_EPS = 1e-9         # :96
MIN_DEV_WINDOW = 5          # :99
MIN_TIME_DIFF = 1e-6        # :1468
SLOPE_WINDOW_SEC = 20       # :1552, :1576
HR_PROMINENCE = 5           # :1496
HR_MIN_DURATION_SEC = 10    # :1486
HR_MIN_CHANGE_BPM = 15      # :1486


def default_params() -> Dict[str, object]:
    """A fresh copy of the hot-path defaults (front end + classifier keys)."""
    return {**HOT_PATH_DEFAULTS, **CLASSIFIER_DEFAULTS}


def band_edges(params: Dict) -> tuple:
    return float(params.get("lowcut_hz", LOWCUT_HZ)), float(params.get("highcut_hz", HIGHCUT_HZ))


def filter_mode(params: Dict) -> str:
    mode = params.get("filter_mode", "parity")
    if mode not in ("parity", "fullrate"):
        raise ValueError(f"filter_mode must be 'parity' or 'fullrate', got {mode!r}")
    return mode


def output_dtype(params: Dict) -> str:
    dt = params.get("output_dtype", "float64")
    if dt not in ("float64", "float32"):
        raise ValueError(f"output_dtype must be 'float64' or 'float32', got {dt!r}")
    return dt


def effective_decimation(sample_rate: int, params: Dict):
    """``(ds, new_rate, clamped)`` as ``bpm_analysis.py:1021-1036`` computes them."""
    ds = params["downsample_factor"]
    _, highcut = band_edges(params)
    max_safe = int((sample_rate / (highcut * 2)) - 1)
    clamped = False
    if ds > max_safe:
        ds = max(1, max_safe)
        clamped = True
    if ds > 1:
        return int(ds), int(sample_rate // ds), clamped
    return 1, int(sample_rate), clamped
