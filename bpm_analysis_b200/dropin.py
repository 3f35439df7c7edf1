"""Device-resident sessions behind the drop-in functions of ``frontend.py``.

``analyze_wav_file`` (bpm_analysis.py:1725-1768) calls the front end as ten separate functions
that hand the SAME host arrays from one to the next: the envelope returned by
``preprocess_audio`` goes to ``_calculate_dynamic_noise_floor`` and to both ``PeakClassifier``
constructions (:1731-1732, :1635, :1740), the floor returned by the second goes to both
``_find_raw_peaks`` calls, the smoothed BPM series to five reductions (:1704-1710).  Served one
call at a time that is an upload, a handful of launches and a blocking read-back per call --
35 ms per 60-minute recording for 0.5 ms of kernels.  Here

  * ``preprocess`` moves only the kept frames x[::ds] to the device (packed by the host cores into
    a pinned staging buffer, ``bpm_host_gather_frames``), enqueues a1..a4 as ONE ``bpm_stage_a`` call
    and reads every result back into pinned host memory with one synchronisation;
  * the arrays it hands out are remembered by identity (address, length, a weak reference that
    proves the memory has not been recycled, and a sampled fingerprint that catches in-place
    edits): when the reference passes them on, the later calls are answered from the session
    without touching the GPU again -- including the second classifier construction;
  * ``beat_metrics`` does the same for a beat list: one upload, a5..a8 in one go, one read-back;
    ``calculate_bpm_series`` registers the Series it returns and the five reductions that take
    that Series look their answer up.

Arrays that did not come out of a session (a caller's own envelope, the reference's envelope in
the parity tests) take the generic path: upload once per session, stay on the device for the
calls that follow.  Nothing here computes on samples with numpy: the host only packs, copies and
bookkeeps.  No CPU fallback: without the libraries or a CUDA device these raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import time
import weakref
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _native as nat
from . import runtime as rt
from .design import PADLEN
from .params import output_dtype

MAX_SESSIONS = 8
MAX_RUNNERS = 4
HOST_GATHER_MIN_PITCH_BYTES = 32      # below this the whole recording is copied (every cache line is touched anyway)


TRACE = os.environ.get("BPM_DROPIN_TRACE", "0") == "1"


class _Trace:
    """Optional host-clock breakdown of the calls (BPM_DROPIN_TRACE=1): name -> accumulated seconds."""

    def __init__(self, sink: Dict[str, float]):
        self.sink, self.t = sink, time.perf_counter()

    def mark(self, name: str) -> None:
        if TRACE:
            now = time.perf_counter()
            self.sink[name] = self.sink.get(name, 0.0) + (now - self.t)
            self.t = now


def _key(a: np.ndarray) -> Tuple[int, int, str]:
    return (a.__array_interface__["data"][0], int(a.size), a.dtype.str)


def _fingerprint(a: np.ndarray) -> bytes:
    """A few sampled elements: catches an array that was edited in place after we handed it out."""
    n = a.size
    if n == 0:
        return b""
    flat = a.reshape(-1)
    idx = (np.arange(24, dtype=np.int64) * (n - 1)) // 23
    return flat[idx].tobytes()


def _pinned(n: int, dtype) -> torch.Tensor:
    return torch.empty(max(int(n), 1), dtype=dtype, pin_memory=True)


class _Watch:
    """Identity of a host array we handed out (or were handed): address + length + dtype, a weak
    reference to the owning object and a fingerprint of its content."""

    def __init__(self, arr: np.ndarray, owner=None):
        self.key = _key(arr)
        self.ref = weakref.ref(owner if owner is not None else arr)
        self.fp = _fingerprint(arr)

    def matches(self, arr: np.ndarray) -> bool:
        return self.ref() is not None and _key(arr) == self.key and _fingerprint(arr) == self.fp


class Session:
    """What is known about one envelope: where it lives on the device and every result computed
    from it so far, each under the configuration it was computed with."""

    def __init__(self, env: np.ndarray, rate: int):
        self.env_watch = _Watch(env)
        self.rate = int(rate)
        self.m = int(env.size)
        self.env_dev: Optional[torch.Tensor] = None
        self.items = None
        self.items_dev = None
        self.a2: Dict[tuple, dict] = {}       # cfg2 -> {"floor", "troughs", "n_all", "mode", "floor_dev", "floor_watch"}
        self.a3: Dict[tuple, dict] = {}       # (cfg3, floor key) -> {"peaks", "strength", "deviation", "smoothed"}


class _StageARunnerSlot:
    """Preallocated buffers for one recording shape: a pregathered StageARunner, the pinned
    staging buffer the host gather fills, pinned result buffers sized by the list bound."""

    def __init__(self, n_in, sample_rate, params, np_dtype, channels, want_debug, want_filtered, sparse):
        self.runner = rt.StageARunner([n_in], sample_rate, params, np_dtype, channels, want_debug=want_debug,
                                      pregathered=("host" if sparse else False), want_filtered=want_filtered or want_debug)
        A = self.runner
        self.sparse = sparse
        n_stage = A.total_m * channels if sparse else n_in * channels
        self.stage = _pinned(n_stage, rt._torch_dtype(np_dtype))
        self.cap = min(A.total_m, A.total_m // max(int(A.cfg.distance), 1) + 2)    # find_peaks distance bounds the lists
        self.f32 = None                               # float32 copies of the signals, allocated on first use
        self.uses = 0
        self.graph = None                             # CUDA graph of the stage-A launches, captured on the second use


def _cfg2(params: Dict, rate: int) -> tuple:
    window = int(params["noise_window_sec"] * rate)                              # bpm_analysis.py:1083
    if window < 3:
        raise ValueError(f"min_periods 3 must be <= window {window}")            # what pandas raises
    distance = int(params["min_peak_distance_sec"] * rate)                       # :1066
    if distance < 1:
        raise ValueError("`distance` must be greater or equal to 1")             # what scipy raises
    return (distance, float(params["trough_prominence_quantile"]), float(params["noise_floor_quantile"]), window,
            float(params.get("trough_rejection_multiplier", 4.0)))


def _cfg3(params: Dict, rate: int) -> tuple:
    distance = int(params["min_peak_distance_sec"] * rate)                       # :226
    if distance < 1:
        raise ValueError("`distance` must be greater or equal to 1")
    return (distance, float(params["peak_prominence_quantile"]), float(params["deviation_smoothing_factor"]))


class DropIn:
    """Process-wide (one per device) service object behind ``frontend``; calls are serialised by a
    lock -- the GPU work of one recording is a single short burst, there is nothing to overlap
    inside one caller, and callers from several threads (gui.py:181-183, Gradio workers) stay safe."""

    def __init__(self):
        self.device = rt.require_cuda()
        self.lib = nat.load_library()
        from . import classifier
        self.host = classifier.load_host_library()
        self.lock = threading.RLock()
        self.sessions: "OrderedDict[tuple, Session]" = OrderedDict()
        self.slots: "OrderedDict[tuple, _StageARunnerSlot]" = OrderedDict()
        self.beat_cache: "OrderedDict[tuple, dict]" = OrderedDict()
        self.series_watch: "OrderedDict[tuple, tuple]" = OrderedDict()
        self.stream = torch.cuda.Stream()
        self.stats = {"session_hits": 0, "session_misses": 0, "stage_a_calls": 0, "beat_hits": 0, "beat_misses": 0}
        self.trace: Dict[str, float] = {}

    def forget(self) -> None:
        """Drop every session and cached result (the preallocated runner buffers stay)."""
        with self.lock:
            self.sessions.clear()
            self.beat_cache.clear()
            from . import frontend
            frontend._series_results.clear()

    # ------------------------------------------------------------------ sessions
    def _remember(self, s: Session) -> None:
        self.sessions[s.env_watch.key] = s
        self.sessions.move_to_end(s.env_watch.key)
        while len(self.sessions) > MAX_SESSIONS:
            self.sessions.popitem(last=False)

    def _session_of(self, env: np.ndarray) -> Optional[Session]:
        s = self.sessions.get(_key(env))
        if s is None:
            return None
        if not s.env_watch.matches(env):
            del self.sessions[s.env_watch.key]
            return None
        self.sessions.move_to_end(s.env_watch.key)
        return s

    def _session_for(self, env: np.ndarray, rate: int) -> Session:
        """The session of this envelope, created (with the envelope uploaded) when it is new."""
        s = self._session_of(env)
        if s is not None and s.rate != int(rate):
            s = None
        if s is None:
            self.stats["session_misses"] += 1
            s = Session(env, rate)
            self._remember(s)
        else:
            self.stats["session_hits"] += 1
        if s.env_dev is None:
            s.env_dev = self._upload(env if env.dtype == np.float64 else env.astype(np.float64))
            s.items = rt.make_items([s.m], [s.m])
            s.items_dev = torch.from_numpy(s.items.view(np.int64).reshape(-1, 4).copy()).to(self.device)
        return s

    def _upload(self, a: np.ndarray) -> torch.Tensor:
        """Pageable host array -> device, through a pinned staging buffer, on our stream."""
        a = np.ascontiguousarray(a)
        stage = _pinned(a.size, rt._torch_dtype(a.dtype))
        np.copyto(stage.numpy()[:a.size], a.reshape(-1))
        with torch.cuda.stream(self.stream):
            dev = torch.empty(max(a.size, 1), dtype=stage.dtype, device=self.device)
            dev[:a.size].copy_(stage[:a.size], non_blocking=True)
        self._keep_alive = stage                      # until the next synchronisation of the stream
        return dev[:a.size] if a.size else dev[:0]

    # ------------------------------------------------------------------ a1 (+ a2..a4 ahead of the calls)
    def _slot(self, n_in, sample_rate, params, np_dtype, channels, want_debug, want_filtered, sparse):
        plan = rt.plan_filter(sample_rate, params)
        key = (int(n_in), int(sample_rate), np.dtype(np_dtype).str, int(channels), bool(want_debug),
               bool(want_filtered), bool(sparse), plan.stride, plan.block, float(plan.low), float(plan.high),
               _cfg2(params, plan.rate), _cfg3(params, plan.rate))
        slot = self.slots.get(key)
        if slot is None:
            slot = _StageARunnerSlot(n_in, sample_rate, params, np_dtype, channels, want_debug, want_filtered, sparse)
            self.slots[key] = slot
            while len(self.slots) > MAX_RUNNERS:
                self.slots.popitem(last=False)
        else:
            self.slots.move_to_end(key)
        return slot

    def preprocess(self, pcm: np.ndarray, sample_rate: int, params: Dict, want_debug: bool, want_filtered: bool):
        """K0+K1+K2(+K2b) of one recording, with a2..a4 computed in the same stage-A call.

        Returns (envelope, rate, filtered | None, debug_int16 | None).  With
        ``params["output_dtype"] == "float32"`` the envelope, the filtered signal and (later, out of the
        session) the noise floor and the per-peak series come back as float32."""
        with self.lock:
            f32 = output_dtype(params) == "float32"
            sig_dtype = torch.float32 if f32 else torch.float64
            from .wav24 import S24Recording
            s24 = pcm if isinstance(pcm, S24Recording) else None      # 24-bit file left in its mapping (wav24.py)
            if s24 is None:
                pcm = np.asarray(pcm)
                if pcm.dtype not in nat.PCM_DTYPES:
                    pcm = pcm.astype(np.float64)
                pcm = np.ascontiguousarray(pcm)
            channels = 1 if pcm.ndim == 1 else int(pcm.shape[1])
            n_in = int(pcm.shape[0])
            plan = rt.plan_filter(sample_rate, params)
            if plan.n_dec(n_in) <= PADLEN:
                raise ValueError("The length of the input vector x must be greater than padlen, which is 15.")
            # bytes between kept frames IN THE SOURCE (3 per sample for a mapped 24-bit file, staged as int32)
            frame_bytes = channels * (3 if s24 is not None else pcm.dtype.itemsize)
            sparse = plan.block == 1 and plan.stride > 1 and plan.stride * frame_bytes >= HOST_GATHER_MIN_PITCH_BYTES
            # the band-passed signal only stays out of HBM in the decimate-first order with a window the fused
            # epilogue takes; otherwise the runner writes it and it is simply not read back unless asked for
            need_filtered = want_filtered or plan.block != 1 or plan.rate // 10 > 65
            tr = _Trace(self.trace)
            slot = self._slot(n_in, sample_rate, params, pcm.dtype, channels, want_debug, need_filtered, sparse)
            A = slot.runner
            M = A.total_m
            tr.mark("pre.plan+slot")
            # -- ingest: the kept frames (or the whole recording) -> pinned staging -> device
            if s24 is not None:
                s24.gather_into(slot.stage.data_ptr(), plan.stride if sparse else 1)
            elif sparse:
                rc = self.host.bpm_host_gather_frames(C.c_void_p(pcm.ctypes.data), frame_bytes, n_in, plan.stride,
                                                      C.c_void_p(slot.stage.data_ptr()), 0)
                if rc != 0:
                    raise RuntimeError(f"bpm_host_gather_frames failed ({rc})")
            else:
                np.copyto(slot.stage.numpy(), pcm.reshape(-1))
            tr.mark("pre.host_gather")
            ev_env, ev_all = torch.cuda.Event(), torch.cuda.Event()
            with torch.cuda.stream(self.stream):
                A.ingest(slot.stage)
                if slot.graph is not None:
                    slot.graph.replay()
                else:
                    A.launch()
                self.stats["stage_a_calls"] += 1
                tr.mark("pre.enqueue_stage_a")
                # result buffers are allocated while the device is already working
                out = {"envelope": _pinned(M, sig_dtype), "floor": _pinned(M, sig_dtype),
                       "counts": _pinned(4, torch.int64)}
                lists = {k: _pinned(slot.cap, torch.int64 if k in ("troughs", "peaks") else sig_dtype)
                         for k in ("troughs", "peaks", "strength", "deviation", "smoothed_dev")}
                extra = {}
                if want_filtered:
                    extra["filtered"] = _pinned(M, sig_dtype)
                if want_debug:
                    extra["debug_wav"] = _pinned(M, torch.int16)
                tr.mark("pre.alloc_results")
                src = A.out if not f32 else self._cast_outputs(slot, M, want_filtered)
                out["envelope"].copy_(src["envelope"][:M], non_blocking=True)
                for k, h in extra.items():
                    h.copy_(src[k][:M] if k == "filtered" else A.out[k], non_blocking=True)
                ev_env.record(self.stream)
                out["floor"].copy_(src["floor"][:M], non_blocking=True)
                out["counts"][0:1].copy_(A.out["trough_count"], non_blocking=True)
                out["counts"][1:2].copy_(A.out["peak_count"], non_blocking=True)
                out["counts"][2:3].copy_(A.out["trough_total"], non_blocking=True)
                out["counts"][3:4].copy_(A.out["floor_mode"], non_blocking=True)
                for k, h in lists.items():
                    h.copy_(src[k][:slot.cap], non_blocking=True)
                ev_all.record(self.stream)
            tr.mark("pre.enqueue_readback")
            ev_env.synchronize()
            tr.mark("pre.wait_envelope")
            env = out["envelope"].numpy()[:M]
            s = Session(env, plan.rate)
            s.pending = (ev_all, out, lists, _cfg2(params, plan.rate), _cfg3(params, plan.rate))
            self._remember(s)
            filt = extra["filtered"].numpy()[:M] if want_filtered else None
            dbg = extra["debug_wav"].numpy()[:M] if want_debug else None
            # a recording shape seen twice gets its stage-A launches captured into a CUDA graph: the next call
            # replays it (one driver call instead of ~30 launches)
            slot.uses += 1
            if slot.uses == 2 and slot.graph is None and not f32 and os.environ.get("BPM_DROPIN_GRAPH", "1") != "0":
                self.stream.synchronize()
                with torch.cuda.stream(self.stream):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        A.launch()
                slot.graph = g
            tr.mark("pre.session")
            return env, plan.rate, filt, dbg

    def _cast_outputs(self, slot: "_StageARunnerSlot", M: int, want_filtered: bool) -> Dict[str, torch.Tensor]:
        """float32 copies (bpm_cast_f32) of the signals of the slot's last stage-A call, on our stream."""
        A = slot.runner
        if slot.f32 is None:
            f = dict(dtype=torch.float32, device=self.device)
            slot.f32 = {"envelope": torch.empty(M, **f), "floor": torch.empty(M, **f), "filtered": torch.empty(M, **f),
                        "strength": torch.empty(slot.cap, **f), "deviation": torch.empty(slot.cap, **f),
                        "smoothed_dev": torch.empty(slot.cap, **f)}
        names = ["envelope", "floor", "strength", "deviation", "smoothed_dev"] + (["filtered"] if want_filtered else [])
        for k in names:
            dst = slot.f32[k]
            nat.check(self.lib.bpm_cast_f32(rt._ptr(A.out[k]), rt._ptr(dst), dst.numel(), self.stream.cuda_stream))
        out = dict(A.out)
        out.update({k: slot.f32[k] for k in names})
        return out

    def _settle(self, s: Session) -> None:
        """Turn the read-back a ``preprocess`` call left in flight into session results."""
        pend = getattr(s, "pending", None)
        if pend is None:
            return
        ev_all, out, lists, cfg2, cfg3 = pend
        ev_all.synchronize()
        s.pending = None
        nt, npk, n_all, mode = (int(v) for v in out["counts"].numpy()[:4])
        floor = out["floor"].numpy()[:s.m]
        troughs = lists["troughs"].numpy()[:nt].copy()
        s.a2[cfg2] = {"floor": floor, "troughs": troughs, "n_all": n_all, "mode": mode, "floor_dev": None,
                      "floor_watch": _Watch(floor)}
        d = max(npk - 1, 0)
        s.a3[(cfg3, _key(floor))] = {"peaks": lists["peaks"].numpy()[:npk].copy(),
                                      "strength": lists["strength"].numpy()[:npk].copy(),
                                      "deviation": lists["deviation"].numpy()[:d].copy(),
                                      "smoothed": lists["smoothed_dev"].numpy()[:d].copy()}

    # ------------------------------------------------------------------ a2
    def noise_floor(self, env: np.ndarray, rate: int, params: Dict):
        """-> (floor float64[M], kept troughs int64[T'], troughs before sanitisation, mode) with
        mode 0 = sanitised floor, 1 = draft floor (<= 2 troughs kept), 2 = static floor (< 5 troughs)."""
        with self.lock:
            cfg = _cfg2(params, rate)
            s = self._session_of(env)
            if s is not None and s.rate == int(rate):
                self._settle(s)
                hit = s.a2.get(cfg)
                if hit is not None and hit["floor_watch"].ref() is not None:
                    self.stats["session_hits"] += 1
                    return hit["floor"], hit["troughs"].copy(), hit["n_all"], hit["mode"]
            s = self._session_for(env, rate)
            n = s.m
            distance, tq, fq, window, mult = cfg
            f_host, c_host = _pinned(n, torch.float64), _pinned(3, torch.int64)
            with torch.cuda.stream(self.stream):
                floor = torch.empty(max(n, 1), dtype=torch.float64, device=self.device)
                tr = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
                cnt = torch.empty(3, dtype=torch.int64, device=self.device)
                nb = int(self.lib.bpm_noise_floor_workspace_bytes(n, 1))
                ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=self.device)
                nat.check(self.lib.bpm_noise_floor(rt._ptr(s.env_dev), rt._ptr(s.items_dev), rt._host_ptr(s.items), 1,
                                                   distance, tq, fq, window, mult, rt._ptr(floor), rt._ptr(tr),
                                                   C.c_void_p(cnt.data_ptr()), C.c_void_p(cnt.data_ptr() + 8),
                                                   C.c_void_p(cnt.data_ptr() + 16), rt._ptr(ws), nb,
                                                   self.stream.cuda_stream))
                f_host[:n].copy_(floor[:n], non_blocking=True)
                c_host.copy_(cnt, non_blocking=True)
                cap = min(n, n // distance + 2)
                t_host = _pinned(cap, torch.int64)
                t_host[:cap].copy_(tr[:cap], non_blocking=True)
            self.stream.synchronize()
            nt, n_all, mode = int(c_host[0]), int(c_host[1]), int(c_host[2])
            fl = f_host.numpy()[:n]
            troughs = t_host.numpy()[:nt].copy()
            s.a2[cfg] = {"floor": fl, "troughs": troughs, "n_all": n_all, "mode": mode, "floor_dev": floor[:n],
                         "floor_watch": _Watch(fl)}
            return fl, troughs.copy(), n_all, mode

    # ------------------------------------------------------------------ a3 / a4
    def raw_peaks(self, env: np.ndarray, floor: np.ndarray, rate: int, params: Dict, with_metrics: bool):
        with self.lock:
            cfg = _cfg3(params, rate)
            s = self._session_of(env)
            if s is not None and s.rate == int(rate):
                self._settle(s)
                hit = s.a3.get((cfg, _key(floor)))
                if hit is not None and self._floor_is_known(s, floor):
                    self.stats["session_hits"] += 1
                    return hit["peaks"].copy(), (hit if with_metrics else None)
            s = self._session_for(env, rate)
            n = s.m
            floor_dev = None
            for e in s.a2.values():                   # a floor this session produced and still holds on the device
                if e["floor_dev"] is not None and e["floor_watch"].matches(floor):
                    floor_dev = e["floor_dev"]
            if floor_dev is None:
                floor_dev = self._upload(floor if floor.dtype == np.float64 else floor.astype(np.float64))
            distance, pq, smoothing = cfg
            cap = min(n, n // distance + 2)
            with torch.cuda.stream(self.stream):
                pk = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
                cnt = torch.empty(1, dtype=torch.int64, device=self.device)
                nb = int(self.lib.bpm_raw_peaks_workspace_bytes(n, 1))
                ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=self.device)
                nat.check(self.lib.bpm_raw_peaks(rt._ptr(s.env_dev), rt._ptr(floor_dev), rt._ptr(s.items_dev),
                                                 rt._host_ptr(s.items), 1, distance, pq, rt._ptr(pk), rt._ptr(cnt),
                                                 rt._ptr(ws), nb, self.stream.cuda_stream))
                f64 = dict(dtype=torch.float64, device=self.device)
                st, dv, sm = torch.empty(max(n, 1), **f64), torch.empty(max(n, 1), **f64), torch.empty(max(n, 1), **f64)
                nat.check(self.lib.bpm_peak_metrics(rt._ptr(s.env_dev), rt._ptr(floor_dev), rt._ptr(pk), rt._ptr(cnt),
                                                    rt._ptr(s.items_dev), rt._host_ptr(s.items), 1, smoothing,
                                                    rt._ptr(st), rt._ptr(dv), rt._ptr(sm), self.stream.cuda_stream))
                c_host = _pinned(1, torch.int64)
                c_host.copy_(cnt, non_blocking=True)
                hp = _pinned(cap, torch.int64)
                hp[:cap].copy_(pk[:cap], non_blocking=True)
                hs, hd, hm = (_pinned(cap, torch.float64) for _ in range(3))
                for h, d in ((hs, st), (hd, dv), (hm, sm)):
                    h[:cap].copy_(d[:cap], non_blocking=True)
            self.stream.synchronize()
            c = int(c_host[0])
            dd = max(c - 1, 0)
            res = {"peaks": hp.numpy()[:c].copy(), "strength": hs.numpy()[:c].copy(),
                   "deviation": hd.numpy()[:dd].copy(), "smoothed": hm.numpy()[:dd].copy()}
            s.a3[(cfg, _key(floor))] = res
            s.floor_watches = getattr(s, "floor_watches", {})
            s.floor_watches[_key(floor)] = _Watch(floor)
            return res["peaks"].copy(), (res if with_metrics else None)

    def _floor_is_known(self, s: Session, floor: np.ndarray) -> bool:
        k = _key(floor)
        for e in s.a2.values():
            if e["floor_watch"].key == k:
                return e["floor_watch"].matches(floor)
        w = getattr(s, "floor_watches", {}).get(k)
        return w is not None and w.matches(floor)

    def note_floor_alias(self, env: np.ndarray, floor: np.ndarray, alias: np.ndarray, owner) -> None:
        """``alias`` (e.g. ``Series.values``) is the same floor held by ``owner``: answer for it too."""
        with self.lock:
            s = self._session_of(env)
            if s is None or _key(alias) == _key(floor):
                return
            for (cfg, fk), res in list(s.a3.items()):
                if fk == _key(floor):
                    s.a3[(cfg, _key(alias))] = res
            s.floor_watches = getattr(s, "floor_watches", {})
            s.floor_watches[_key(alias)] = _Watch(alias, owner)

    # ------------------------------------------------------------------ a5..a8
    def beat_metrics(self, beats: np.ndarray, rate: int, params: Dict) -> dict:
        """Everything bpm_analysis.py:1704-1710 derives from one beat list, in one device round trip:
        the BPM series (K9), both steepest slopes (K10), the extrema of the smoothed series for the
        incline / decline search (K11) and the windowed HRV table (K12)."""
        with self.lock:
            beats = np.ascontiguousarray(beats, dtype=np.int64)
            window_us = rt.smoothing_window_us(params)
            win, step = int(params["hrv_window_size_beats"]), int(params["hrv_step_size_beats"])
            key = (hash(beats.tobytes()), beats.size, int(rate), window_us, win, step)
            hit = self.beat_cache.get(key)
            if hit is not None:
                self.beat_cache.move_to_end(key)
                self.stats["beat_hits"] += 1
                return hit
            self.stats["beat_misses"] += 1
            res = rt.beat_chain(self, beats, int(rate), window_us, win, step)
            self.beat_cache[key] = res
            while len(self.beat_cache) > 16:
                self.beat_cache.popitem(last=False)
            return res


_singleton: Dict[int, DropIn] = {}
_singleton_lock = threading.Lock()


def dropin() -> DropIn:
    rt.require_cuda()
    dev = torch.cuda.current_device()
    with _singleton_lock:
        d = _singleton.get(dev)
        if d is None:
            d = _singleton[dev] = DropIn()
        return d
