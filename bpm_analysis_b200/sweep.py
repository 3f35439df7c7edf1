"""Parameter sweep over ONE recording (BASELINE configs[4], "C5"): many band-pass / noise-floor
settings, the recording resident on the GPU once.

Settings are independent units of work, so a multi-GPU sweep replicates the recording and
shards the SETTINGS over the ranks with no collective (first row); within a rank
the settings are grouped by band-pass so that the filter + envelope (a1) runs once per distinct
(lowcut, highcut) and only the noise-floor / peak stages (a2..a4) run per setting.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .dist import shard_range
from .stream import DeviceEngine


def run_sweep(pcm: np.ndarray, sample_rate: int, base_params: Dict, settings: Sequence[Dict], rank: int = 0,
              world: int = 1, keep_arrays: bool = False, engine: Optional[DeviceEngine] = None) -> List[Dict]:
    """a1..a4 for every setting of this rank's block of ``settings`` (dicts of parameter
    overrides).  Returns one dict per setting: index, rate, m, trough / peak counts and, with
    ``keep_arrays``, the device tensors (envelope, floor, troughs, peaks, strength, smoothed_dev);
    a setting whose band-pass the reference rejects (ValueError, :1041-1042) yields
    ``{"setting": i, "error": message}``."""
    from .runtime import plan_filter, to_device
    E = engine or DeviceEngine()
    lo, hi = shard_range(len(settings), world, rank)
    mine = list(range(lo, hi))
    pcm = np.asarray(pcm)
    channels = 1 if pcm.ndim == 1 else int(pcm.shape[1])
    pcm_dev = to_device(pcm.reshape(-1))
    groups: Dict[tuple, List[int]] = {}
    for i in mine:
        p = dict(base_params, **settings[i])
        key = (float(p.get("lowcut_hz", 20.0)), float(p.get("highcut_hz", 150.0)), int(p["downsample_factor"]),
               str(p.get("filter_mode", "parity")))
        groups.setdefault(key, []).append(i)
    out: Dict[int, Dict] = {}
    for key, idxs in groups.items():
        p0 = dict(base_params, **settings[idxs[0]])
        try:
            plan = plan_filter(sample_rate, p0)
        except ValueError as e:
            # the reference raises for this band-pass (bpm_analysis.py:1041-1042); its callers catch
            # per file, a sweep records the error per setting and goes on
            for i in idxs:
                out[i] = {"setting": i, "error": str(e)}
            continue
        _, env = E.frontend(pcm_dev, int(pcm.shape[0]), plan, channels, pcm.dtype)
        rate = plan.rate
        for i in idxs:
            p = dict(base_params, **settings[i])
            window = int(p["noise_window_sec"] * rate)                       # bpm_analysis.py:1083
            if window < 3:
                raise ValueError(f"min_periods 3 must be <= window {window}")
            distance = int(p["min_peak_distance_sec"] * rate)               # :226, :1066
            if distance < 1:
                raise ValueError("`distance` must be greater or equal to 1")
            floor, troughs = E.noise_floor(env, distance, window, p)
            peaks = E.raw_peaks(env, floor, distance, float(p["peak_prominence_quantile"]))
            strength, deviation, smoothed = E.peak_metrics(env, floor, peaks, float(p["deviation_smoothing_factor"]))
            rec = {"setting": i, "rate": rate, "m": int(env.numel()), "n_troughs": int(troughs.numel()),
                   "n_peaks": int(peaks.numel())}
            if keep_arrays:
                rec.update(envelope=env, floor=floor, troughs=troughs, peaks=peaks, strength=strength,
                           smoothed_dev=smoothed)
            out[i] = rec
    torch.cuda.synchronize()
    return [out[i] for i in mine]
