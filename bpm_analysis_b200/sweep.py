"""Parameter sweep over ONE recording (BASELINE configs[4], "C5"): many band-pass / noise-floor
settings, the recording resident on the GPU once.

Settings are independent units of work, so a multi-GPU sweep replicates the recording and
shards the SETTINGS over the ranks with no collective (first row).  Within a rank
the settings are grouped by band-pass so that the filter + envelope (a1) runs once per distinct
(lowcut, highcut); the noise-floor / peak / metric stages (a2..a4) of the settings are then
enqueued round-robin on a few CUDA streams -- every stage is latency-bound at this size, so
independent settings overlap on the device -- with all counts left on the device: the host
synchronises ONCE, at the end of the sweep.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as nat
from . import runtime as rt
from .dist import shard_range

N_STREAMS = 4


class _Lane:
    """One stream with its own workspace (calls on different streams must not share scratch)."""

    def __init__(self, lib, device, m: int):
        self.stream = torch.cuda.Stream()
        nb = max(int(lib.bpm_noise_floor_workspace_bytes(m, 1)), int(lib.bpm_raw_peaks_workspace_bytes(m, 1)), 256)
        self.ws = torch.empty(nb, dtype=torch.uint8, device=device)
        self.nb = nb


def run_sweep(pcm: np.ndarray, sample_rate: int, base_params: Dict, settings: Sequence[Dict], rank: int = 0,
              world: int = 1, keep_arrays: bool = False, engine=None, n_streams: int = N_STREAMS) -> List[Dict]:
    """a1..a4 for every setting of this rank's block of ``settings`` (dicts of parameter
    overrides).  Returns one dict per setting: index, rate, m, trough / peak counts and, with
    ``keep_arrays``, the device tensors (envelope, floor, troughs, peaks, strength, smoothed_dev);
    a setting whose band-pass the reference rejects (ValueError, :1041-1042) yields
    ``{"setting": i, "error": message}``."""
    device = rt.require_cuda()
    lib = nat.load_library()
    lo, hi = shard_range(len(settings), world, rank)
    mine = list(range(lo, hi))
    pcm = np.asarray(pcm)
    channels = 1 if pcm.ndim == 1 else int(pcm.shape[1])
    n_in = int(pcm.shape[0])
    pcm_dev = rt.to_device(pcm.reshape(-1))
    groups: Dict[tuple, List[int]] = {}
    for i in mine:
        p = dict(base_params, **settings[i])
        key = (float(p.get("lowcut_hz", 20.0)), float(p.get("highcut_hz", 150.0)), int(p["downsample_factor"]),
               str(p.get("filter_mode", "parity")))
        groups.setdefault(key, []).append(i)
    out: Dict[int, Dict] = {}
    main = torch.cuda.current_stream()
    lanes: Dict[int, List[_Lane]] = {}
    pending = []                                   # (setting, record, counts tensor) resolved after the one sync
    keep_alive = []
    f64 = dict(dtype=torch.float64, device=device)
    i64 = dict(dtype=torch.int64, device=device)
    turn = 0
    for key, idxs in groups.items():
        p0 = dict(base_params, **settings[idxs[0]])
        try:
            plan = rt.plan_filter(sample_rate, p0)
        except ValueError as e:
            # the reference raises for this band-pass (bpm_analysis.py:1041-1042); its callers catch
            # per file, a sweep records the error per setting and goes on
            for i in idxs:
                out[i] = {"setting": i, "error": str(e)}
            continue
        m, rate = plan.m(n_in), plan.rate
        items = rt.make_items([n_in], [m])
        items_dev = torch.from_numpy(items.view(np.int64).reshape(-1, 4).copy()).to(device)
        eitems = rt.make_items([m], [m])
        eitems_dev = torch.from_numpy(eitems.view(np.int64).reshape(-1, 4).copy()).to(device)
        design, design_host = rt.design_images(plan)
        env, amax = torch.empty(m, **f64), torch.empty(1, **f64)
        nb = int(lib.bpm_frontend_workspace_bytes(m, 1))
        ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=device)
        fused = plan.block == 1 and rate // 10 <= 65
        filt = None if fused else torch.empty(m, **f64)
        nat.check(lib.bpm_frontend(rt._ptr(pcm_dev), nat.PCM_DTYPES[pcm.dtype], channels, rt._ptr(items_dev),
                                   rt._host_ptr(items), 1, plan.stride, rt._ptr(design), rt._host_ptr(design_host),
                                   int(design.numel()), rate // 10, rt._ptr(filt), rt._ptr(env), rt._ptr(amax),
                                   rt._ptr(ws), nb, main.cuda_stream))
        ready = torch.cuda.Event()
        ready.record(main)
        keep_alive += [ws, filt, items_dev, eitems_dev, amax]
        if m not in lanes:
            lanes[m] = [_Lane(lib, device, m) for _ in range(max(1, n_streams))]
        for i in idxs:
            p = dict(base_params, **settings[i])
            window = int(p["noise_window_sec"] * rate)                       # bpm_analysis.py:1083
            if window < 3:
                raise ValueError(f"min_periods 3 must be <= window {window}")
            distance = int(p["min_peak_distance_sec"] * rate)               # :226, :1066
            if distance < 1:
                raise ValueError("`distance` must be greater or equal to 1")
            lane = lanes[m][turn % len(lanes[m])]
            turn += 1
            floor, tr, pk = torch.empty(m, **f64), torch.empty(m, **i64), torch.empty(m, **i64)
            st, dv, sm = torch.empty(m, **f64), torch.empty(m, **f64), torch.empty(m, **f64)
            cnt = torch.empty(4, **i64)                                      # kept troughs, all troughs, mode, raw peaks
            base = cnt.data_ptr()
            with torch.cuda.stream(lane.stream):
                lane.stream.wait_event(ready)
                s = lane.stream.cuda_stream
                nat.check(lib.bpm_noise_floor(rt._ptr(env), rt._ptr(eitems_dev), rt._host_ptr(eitems), 1, distance,
                                              float(p["trough_prominence_quantile"]), float(p["noise_floor_quantile"]),
                                              window, float(p.get("trough_rejection_multiplier", 4.0)), rt._ptr(floor),
                                              rt._ptr(tr), C.c_void_p(base), C.c_void_p(base + 8), C.c_void_p(base + 16),
                                              rt._ptr(lane.ws), lane.nb, s))
                nat.check(lib.bpm_raw_peaks(rt._ptr(env), rt._ptr(floor), rt._ptr(eitems_dev), rt._host_ptr(eitems), 1,
                                            distance, float(p["peak_prominence_quantile"]), rt._ptr(pk),
                                            C.c_void_p(base + 24), rt._ptr(lane.ws), lane.nb, s))
                nat.check(lib.bpm_peak_metrics(rt._ptr(env), rt._ptr(floor), rt._ptr(pk), C.c_void_p(base + 24),
                                               rt._ptr(eitems_dev), rt._host_ptr(eitems), 1,
                                               float(p["deviation_smoothing_factor"]), rt._ptr(st), rt._ptr(dv),
                                               rt._ptr(sm), s))
            rec = {"setting": i, "rate": rate, "m": m}
            arrays = dict(envelope=env, floor=floor, troughs=tr, peaks=pk, strength=st, smoothed_dev=sm)
            pending.append((rec, cnt, arrays if keep_arrays else None))
            if not keep_arrays:
                keep_alive += [floor, tr, pk, st, dv, sm]
            out[i] = rec
    for ls in lanes.values():
        for lane in ls:
            main.wait_stream(lane.stream)
    torch.cuda.synchronize()                       # the ONE host synchronisation of the sweep
    if pending:
        counts = torch.stack([c for _, c, _ in pending]).cpu().numpy()
        for (rec, _, arrays), c in zip(pending, counts):
            nt, npk = int(c[0]), int(c[3])
            rec["n_troughs"], rec["n_peaks"] = nt, npk
            if arrays is not None:
                d = max(npk - 1, 0)
                rec.update(envelope=arrays["envelope"], floor=arrays["floor"], troughs=arrays["troughs"][:nt],
                           peaks=arrays["peaks"][:npk], strength=arrays["strength"][:npk],
                           smoothed_dev=arrays["smoothed_dev"][:d])
    del keep_alive
    return [out[i] for i in mine]
