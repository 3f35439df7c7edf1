"""Synthetic heart-sound (PCG) recordings for the BASELINE.json configs.

Generator spec: SURVEY.md §8(d).  Each beat is an S1 burst (Gaussian-windowed
50 Hz tone, sigma 25 ms) followed by an S2 burst (70 Hz, sigma 18 ms) on a
white-noise bed, scaled to 95 % of int16 full scale.  Everything is seeded
with ``numpy.random.default_rng`` so the GPU box and the authoring container
produce identical recordings.

The generators return ``(pcm int16[N], sample_rate, beat_times_sec)``; the
beat times are the ground-truth S1 onsets and are what ``bench.py`` feeds the
beat-list reductions (BPM series / slopes / HRV) with, because the sequential
S1/S2 classifier that normally produces the beat list is outside the hot path.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "pcg_recording", "config_c1", "config_c2", "config_c3_item", "config_c4",
    "config_c5", "beats_to_envelope_indices", "CONFIG_SHAPES",
]

# name -> (duration_sec, sample_rate)
CONFIG_SHAPES = {
    "C1": (300.0, 44100),
    "C2": (3600.0, 48000),
    "C3": (600.0, 44100),
    "C4": (86400.0, 4000),
    "C5": (1800.0, 48000),
}


def _add_burst(x: np.ndarray, sr: int, t0: float, freq: float, sigma: float, amp: float) -> None:
    lo = int(np.floor((t0 - 4.0 * sigma) * sr))
    hi = int(np.ceil((t0 + 4.0 * sigma) * sr)) + 1
    lo_c, hi_c = max(lo, 0), min(hi, x.shape[0])
    if hi_c <= lo_c:
        return
    tau = np.arange(lo_c, hi_c, dtype=np.float64) / sr - t0
    x[lo_c:hi_c] += amp * np.exp(-0.5 * (tau / sigma) ** 2) * np.sin(2.0 * np.pi * freq * tau)


def pcg_recording(duration_sec: float, sample_rate: int, bpm_of_t, seed: int,
                  noise_sigma: float = 0.02, bursts=None, dropouts=None):
    """Build one recording.

    ``bpm_of_t`` maps time (s) -> instantaneous BPM.  ``bursts`` / ``dropouts``
    are optional lists of ``(start_sec, length_sec)``: noise x10 / hard zeros.
    """
    rng = np.random.default_rng(seed)
    n = int(round(duration_sec * sample_rate))
    # noise bed in float32 chunks keeps the 24 h config inside a few GB
    x = np.empty(n, dtype=np.float64)
    step = 1 << 22
    for s in range(0, n, step):
        e = min(n, s + step)
        x[s:e] = rng.standard_normal(e - s) * noise_sigma
    if bursts:
        for (bs, bl) in bursts:
            a, b = int(bs * sample_rate), min(n, int((bs + bl) * sample_rate))
            x[a:b] *= 10.0
    beats = []
    t = 0.5
    while t < duration_sec - 1.0:
        rr = 60.0 / float(bpm_of_t(t))
        beats.append(t)
        a1 = 1.0 * (1.0 + 0.05 * rng.standard_normal())
        a2 = 0.6 * (1.0 + 0.05 * rng.standard_normal())
        _add_burst(x, sample_rate, t, 50.0, 0.025, a1)
        _add_burst(x, sample_rate, t + min(0.30, 0.35 * rr), 70.0, 0.018, a2)
        t += rr * (1.0 + 0.02 * rng.standard_normal())
    if dropouts:
        for (ds_, dl) in dropouts:
            a, b = int(ds_ * sample_rate), min(n, int((ds_ + dl) * sample_rate))
            x[a:b] = 0.0
    peak = np.max(np.abs(x))
    pcm = np.round(x * (0.95 * 32767.0 / peak)).astype(np.int16)
    return pcm, int(sample_rate), np.asarray(beats, dtype=np.float64)


def config_c1(seed: int = 1, duration_sec: float = 300.0):
    """C1: 5 min, 44.1 kHz, 70 BPM."""
    return pcg_recording(duration_sec, 44100, lambda t: 70.0, seed)


def _ramp_profile(duration_sec: float):
    a, b, c = 0.15 * duration_sec, 0.50 * duration_sec, 0.90 * duration_sec

    def bpm(t: float) -> float:
        if t < a:
            return 60.0
        if t < b:
            return 60.0 + (170.0 - 60.0) * (t - a) / (b - a)
        if t < c:
            return 170.0 + (80.0 - 170.0) * (t - b) / (c - b)
        return 80.0
    return bpm


def config_c2(seed: int = 2, duration_sec: float = 3600.0, sample_rate: int = 48000):
    """C2: 60 min, 48 kHz, 60 -> 170 -> 80 BPM exertion/recovery ramp."""
    return pcg_recording(duration_sec, sample_rate, _ramp_profile(duration_sec), seed)


def config_c3_item(item: int, duration_sec: float = 600.0):
    """C3: item ``item`` of the 1024 x 10 min 44.1 kHz batch (seed 1000+item)."""
    seed = 1000 + int(item)
    bpm = float(np.random.default_rng(seed ^ 0x5EED).uniform(55.0, 95.0))
    return pcg_recording(duration_sec, 44100, lambda t: bpm, seed)


def config_c4(seed: int = 4, duration_sec: float = 86400.0):
    """C4: 24 h, 4 kHz Holter-style, slow +-10 BPM drift, noise bursts, dropouts."""
    rng = np.random.default_rng(seed + 7919)
    bursts, dropouts = [], []
    t = rng.exponential(600.0)
    while t < duration_sec:
        bursts.append((t, float(rng.uniform(2.0, 10.0))))
        t += rng.exponential(600.0)
    t = rng.exponential(1200.0)
    while t < duration_sec:
        dropouts.append((t, float(rng.uniform(1.0, 5.0))))
        t += rng.exponential(1200.0)
    period = 1800.0
    return pcg_recording(duration_sec, 4000,
                         lambda tt: 70.0 + 10.0 * np.sin(2.0 * np.pi * tt / period),
                         seed, bursts=bursts, dropouts=dropouts)


def config_c5(seed: int = 5, duration_sec: float = 1800.0):
    """C5: the one 30 min 48 kHz recording shared by the 256-setting sweep."""
    return pcg_recording(duration_sec, 48000, _ramp_profile(duration_sec), seed)


def c5_settings():
    """The 256 sweep settings: 16 band edges x 16 noise-floor settings."""
    out = []
    for lo in (15.0, 20.0, 25.0, 30.0):
        for hi in (100.0, 120.0, 150.0, 200.0):
            for q in (0.1, 0.15, 0.2, 0.3):
                for w in (5, 10, 15, 20):
                    out.append({"lowcut_hz": lo, "highcut_hz": hi,
                                "noise_floor_quantile": q, "noise_window_sec": w})
    return out


def beats_to_envelope_indices(beat_times_sec: np.ndarray, envelope_rate: int) -> np.ndarray:
    """Ground-truth beat onsets -> int64 envelope-sample indices (strictly increasing)."""
    idx = np.round(np.asarray(beat_times_sec) * envelope_rate).astype(np.int64)
    return np.unique(idx)
