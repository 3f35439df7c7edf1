"""Device-side plumbing: batches of recordings -> libbpm_b200 calls.

PyTorch owns device memory, pinned staging buffers and streams; every numeric
step is a call into the C ABI (``_native.py``).  Nothing here computes on
samples with numpy -- the host only builds small descriptor arrays (offsets,
lengths, filter design tables).
"""
from __future__ import annotations

import ctypes as C
import time
import os
import threading
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as nat
from .design import PADLEN, design_block_filter
from .params import band_edges, effective_decimation, filter_mode


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise nat.NativeLibraryError("no CUDA device visible: the front end runs on the GPU only "
                                     "(there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


# Ops.frontend moves only the kept frames when consecutive kept frames are at least this many bytes
# apart in the caller's (pageable) array.  Measured on the B200 box for C2: the driver-staged strided
# copy costs ~11 ns per kept frame, a plain pageable copy ~0.09 ns per byte -> break-even ~128 bytes
# (stride 64 for mono int16; the reference's defaults give 292 / 318 bytes at 44.1 / 48 kHz).  Strides
# under 40 always take the plain copy.
SPARSE_INGEST_MIN_PITCH_BYTES = 128


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _host_ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def make_items(n_in: Sequence[int], m: Sequence[int]) -> np.ndarray:
    """Descriptor array (BpmItem) for recordings packed back to back."""
    n_in = np.asarray(n_in, dtype=np.int64)
    m = np.asarray(m, dtype=np.int64)
    it = np.zeros(len(n_in), dtype=nat.ITEM_DTYPE)
    it["n_in"], it["m"] = n_in, m
    it["in_off"][1:] = np.cumsum(n_in)[:-1]
    it["m_off"][1:] = np.cumsum(m)[:-1]
    return it


def to_device(a: np.ndarray) -> torch.Tensor:
    require_cuda()
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:          # torch refuses read-only views (e.g. arrays out of np.load)
        a = a.copy()
    return torch.from_numpy(a).cuda(non_blocking=False)


@dataclass
class FilterPlan:
    """Where the band-pass runs and what it decimates to (bpm_analysis.py:1018-1045)."""
    sample_rate: int
    ds: int
    rate: int            # envelope rate = sample_rate // ds
    stride: int          # samples skipped before the filter (ds in 'parity', 1 in 'fullrate')
    block: int           # filter outputs skipped (1 in 'parity', ds in 'fullrate')
    clamped: bool
    design: object       # BlockFilterDesign
    low: float = 0.0     # band edges as fractions of Nyquist at the rate the filter runs at
    high: float = 0.0

    def n_dec(self, n_in: int) -> int:
        return (n_in + self.stride - 1) // self.stride

    def m(self, n_in: int) -> int:
        return (self.n_dec(n_in) + self.block - 1) // self.block


def plan_filter(sample_rate: int, params: Dict) -> FilterPlan:
    ds, rate, clamped = effective_decimation(sample_rate, params)
    lowcut, highcut = band_edges(params)
    mode = filter_mode(params)
    if mode == "parity":
        nyq = 0.5 * rate
        lo, hi = lowcut / nyq, highcut / nyq
        if hi >= 1.0:                                           # bpm_analysis.py:1041-1042
            raise ValueError(f"Cannot create a {int(highcut) if float(highcut).is_integer() else highcut}Hz "
                             f"filter. The effective sample rate of {rate}Hz is too low.")
        return FilterPlan(sample_rate, ds, rate, ds, 1, clamped, design_block_filter(lo, hi, 1), lo, hi)
    nyq = 0.5 * sample_rate
    lo, hi = lowcut / nyq, highcut / nyq
    if hi >= 1.0:
        raise ValueError(f"Cannot create a {highcut}Hz filter. The sample rate of {sample_rate}Hz is too low.")
    return FilterPlan(sample_rate, ds, rate, 1, ds, clamped, design_block_filter(lo, hi, ds), lo, hi)


DESIGN_CACHE_SIZE = 64
_design_cache: "OrderedDict[tuple, tuple]" = OrderedDict()
_design_lock = threading.Lock()


def design_images(plan: FilterPlan):
    """(device tensor, host array) of the packed design image of ``plan``.

    Cached by VALUE -- (device, band edges, block) -- and bounded (LRU): a long-lived process or a
    band-edge sweep creates many designs, and an identity key would be reused by CPython after the
    design object is collected.  The host array is what the decimate-first kernels read (their
    coefficients travel as launch parameters); an entry keeps both alive while it is cached."""
    key = (torch.cuda.current_device(), float(plan.low), float(plan.high), int(plan.block))
    with _design_lock:
        hit = _design_cache.get(key)
        if hit is not None:
            _design_cache.move_to_end(key)
            return hit
    host = np.ascontiguousarray(plan.design.packed(), dtype=np.float64)
    entry = (to_device(host), host)
    with _design_lock:
        _design_cache[key] = entry
        while len(_design_cache) > DESIGN_CACHE_SIZE:
            _design_cache.popitem(last=False)
    return entry


def design_on_device(plan: FilterPlan) -> torch.Tensor:
    return design_images(plan)[0]


def stage_a_config(plan: FilterPlan, params: Dict, pcm_dtype: int, channels: int, want_debug: bool
                   ) -> nat.StageAConfig:
    rate = plan.rate
    noise_window = int(params["noise_window_sec"] * rate)                   # :1083
    if noise_window < 3:
        raise ValueError(f"min_periods 3 must be <= window {noise_window}")    # what pandas raises
    distance = int(params["min_peak_distance_sec"] * rate)                  # :226, :1066
    if distance < 1:
        raise ValueError("`distance` must be greater or equal to 1")           # what scipy raises
    return nat.StageAConfig(
        stride=plan.stride, block=plan.block, pcm_dtype=pcm_dtype, channels=channels,
        env_window=rate // 10, distance=distance, noise_window=noise_window,
        want_debug_wav=1 if want_debug else 0,
        trough_prom_q=float(params["trough_prominence_quantile"]),
        peak_prom_q=float(params["peak_prominence_quantile"]),
        floor_q=float(params["noise_floor_quantile"]),
        rejection_multiplier=float(params.get("trough_rejection_multiplier", 4.0)),
        smoothing_factor=float(params["deviation_smoothing_factor"]))


class StageAResult:
    """Device tensors of one stage-A call plus lazy host views per recording."""

    def __init__(self, items: np.ndarray, rate: int, dev: Dict[str, torch.Tensor]):
        self.items, self.rate, self.dev = items, rate, dev
        self._host: Dict[str, np.ndarray] = {}

    def host(self, name: str) -> np.ndarray:
        if name not in self._host:
            self._host[name] = self.dev[name].cpu().numpy()
        return self._host[name]

    def item(self, i: int) -> Dict[str, object]:
        it = self.items[i]
        o, m = int(it["m_off"]), int(it["m"])
        nt, npk = int(self.host("trough_count")[i]), int(self.host("peak_count")[i])
        out = {"rate": self.rate,
               "filtered": self.host("filtered")[o:o + m] if "filtered" in self.dev else None,
               "envelope": self.host("envelope")[o:o + m],
               "floor": self.host("floor")[o:o + m], "absmax": float(self.host("absmax")[i]),
               "troughs": self.host("troughs")[o:o + nt].copy(), "peaks": self.host("peaks")[o:o + npk].copy(),
               "strength": self.host("strength")[o:o + npk],
               "deviation": self.host("deviation")[o:o + max(npk - 1, 0)],
               "smoothed_dev": self.host("smoothed_dev")[o:o + max(npk - 1, 0)]}
        if "debug_wav" in self.dev:
            out["debug_wav"] = self.host("debug_wav")[o:o + m]
        return out


class StageARunner:
    """a1..a4 for a batch of equal-format recordings in one C call (bpm_stage_a).

    Buffers are allocated once for a given batch shape and reused, so repeated
    calls (bench.py, servers) enqueue work without touching the allocator.
    """

    def __init__(self, n_in: Sequence[int], sample_rate: int, params: Dict, pcm_dtype=np.int16,
                 channels: int = 1, want_debug: bool = False, pregathered: bool = False, want_filtered: bool = True):
        """``pregathered``: the decimation x[::ds] (K0) runs on its own (``gather``) into a frame
        buffer that stage A then reads with stride 1 -- bit-identical to the fused path, and it
        lets ``StageAPipeline`` overlap the PCIe-bound ingest of the next recording with the
        compute of the current one.  ``True`` / ``"sm"``: bpm_gather_frames, a small kernel that
        writes float64 frames; ``"ce"``: bpm_copy_frames, one strided 2-D copy on the copy engine
        that keeps the PCM's dtype and channels (no SM involved, faster over PCIe); ``"host"``: the
        frames were packed by the host cores (bpm_host_gather_frames) into a pinned staging buffer
        that ``ingest`` copies in one piece -- same device layout as ``"ce"``.
        Decimate-then-filter order only.
        ``want_filtered=False``: the band-passed signal is not written to HBM at all (the envelope
        is formed inside the backward pass); only the debug WAV and the tests read it."""
        self.device = require_cuda()
        self.lib = nat.load_library()
        self.plan = plan_filter(sample_rate, params)
        self.np_dtype = np.dtype(pcm_dtype)
        if self.np_dtype not in nat.PCM_DTYPES:
            raise TypeError(f"unsupported PCM dtype {self.np_dtype}")
        self.channels = int(channels)
        for n in n_in:
            if self.plan.n_dec(int(n)) <= PADLEN:                     # scipy _validate_pad
                raise ValueError("The length of the input vector x must be greater than padlen, which is 15.")
        self.items = make_items(n_in, [self.plan.m(int(n)) for n in n_in])
        self.n_items = len(self.items)
        self.total_in = int(self.items["n_in"].sum())
        self.total_m = int(self.items["m"].sum())
        self.cfg = stage_a_config(self.plan, params, nat.PCM_DTYPES[self.np_dtype], self.channels, want_debug)
        if pregathered not in (False, True, "sm", "ce", "host"):
            raise ValueError("pregathered must be False, True / 'sm', 'ce' or 'host'")
        self.pregathered = bool(pregathered)
        self.ingest_kind = pregathered if pregathered in ("ce", "host") else ("sm" if pregathered else None)
        if self.pregathered:
            if self.plan.block != 1:
                raise ValueError("pregathered ingest needs the decimate-then-filter order (filter_mode 'parity')")
            self.src_items = self.items
            self.src_items_dev = torch.from_numpy(self.src_items.view(np.int64).reshape(-1, 4).copy()).to(self.device)
            self.src_stride = int(self.plan.stride)
            self.items = make_items(self.src_items["m"], self.src_items["m"])
            self.cfg.stride = 1
            if self.ingest_kind == "sm":
                self.cfg.pcm_dtype, self.cfg.channels = nat.PCM_DTYPES[np.dtype(np.float64)], 1
        self.items_dev = torch.from_numpy(self.items.view(np.int64).reshape(-1, 4).copy()).to(self.device)
        self.design_dev, self.design_host = design_images(self.plan)
        self.design_words = int(self.design_dev.numel())
        f64 = dict(dtype=torch.float64, device=self.device)
        i64 = dict(dtype=torch.int64, device=self.device)
        M, n = self.total_m, self.n_items
        self.out = {"filtered": torch.empty(M, **f64), "envelope": torch.empty(M, **f64),
                    "absmax": torch.empty(n, **f64), "floor": torch.empty(M, **f64),
                    "troughs": torch.empty(M, **i64), "trough_count": torch.empty(n, **i64),
                    "peaks": torch.empty(M, **i64), "peak_count": torch.empty(n, **i64),
                    "strength": torch.empty(M, **f64), "deviation": torch.empty(M, **f64),
                    "smoothed_dev": torch.empty(M, **f64), "trough_total": torch.empty(n, **i64),
                    "floor_mode": torch.empty(n, **i64)}
        self.want_filtered = bool(want_filtered) or bool(want_debug)
        if not self.want_filtered:
            if self.plan.block != 1 or self.cfg.env_window > 65:
                raise ValueError("want_filtered=False needs the decimate-then-filter order and an envelope window <= 65")
            del self.out["filtered"]
        if want_debug:
            self.out["debug_wav"] = torch.empty(M, dtype=torch.int16, device=self.device)
        self.ws_bytes = int(self.lib.bpm_stage_a_workspace_bytes(M, n))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        if self.pregathered:
            self.frames_dev = (torch.empty(M, dtype=torch.float64, device=self.device) if self.ingest_kind == "sm" else
                               torch.empty(M * self.channels, dtype=_torch_dtype(self.np_dtype), device=self.device))
            self.pcm_dev = None
            self._pcm_src = self.frames_dev
        else:
            self.pcm_dev = torch.empty(self.total_in * self.channels, dtype=_torch_dtype(self.np_dtype),
                                       device=self.device)
            self._pcm_src = self.pcm_dev      # what the kernels read: the device copy, or mapped pinned host memory
        self.outs_struct = nat.StageAOutputs(**{k: self.out[k].data_ptr() if k in self.out else None
                                                for k, _ in nat.StageAOutputs._fields_})

    # -- input staging
    def upload(self, pcms: Sequence[np.ndarray]) -> None:
        """Pageable host arrays -> the device PCM buffer (one copy per recording)."""
        off = 0
        for a in pcms:
            a = np.ascontiguousarray(a)
            flat = torch.from_numpy(a.reshape(-1))
            self.pcm_dev[off:off + flat.numel()].copy_(flat, non_blocking=False)
            off += flat.numel()

    def upload_pinned(self, pinned: torch.Tensor) -> None:
        """One async H2D copy from a pinned host tensor holding the whole batch."""
        self.pcm_dev.copy_(pinned, non_blocking=True)
        self._pcm_src = self.pcm_dev

    def read_from_host(self, pinned: torch.Tensor) -> None:
        """Zero-copy ingest: the kernels read the batch straight out of PINNED host memory.

        Worth it in the reference's decimate-then-filter order, where only every ds-th frame is
        ever touched: the sectors holding kept frames cross PCIe, not the whole recording.  The
        tensor must stay alive (and unchanged) until the step has run.
        """
        if not pinned.is_pinned() or pinned.numel() != self.pcm_dev.numel() or pinned.dtype != self.pcm_dev.dtype:
            raise ValueError("need a pinned host tensor with the batch's dtype and size")
        self._pcm_src = pinned

    def gather(self, pcm: torch.Tensor) -> None:
        """K0 on the current stream (pregathered mode): kept frames of ``pcm`` -> the frame buffer.
        ``pcm`` is a device tensor or a PINNED host tensor ('sm': the kernel reads it over PCIe,
        zero-copy; 'ce': the copy engine fetches one frame per row of a strided 2-D copy)."""
        if not self.pregathered:
            raise RuntimeError("runner was not built with pregathered=True")
        if not (pcm.is_cuda or pcm.is_pinned()) or pcm.numel() != self.total_in * self.channels \
                or pcm.dtype != _torch_dtype(self.np_dtype):
            raise ValueError("need a device or pinned host tensor with the batch's dtype and size")
        if self.ingest_kind == "host":
            raise RuntimeError("host-gathered runner: pack the frames with bpm_host_gather_frames and call ingest()")
        if self.ingest_kind == "ce":
            nat.check(self.lib.bpm_copy_frames(_ptr(pcm), nat.PCM_DTYPES[self.np_dtype], self.channels,
                                               _host_ptr(self.src_items), self.n_items, self.src_stride,
                                               _ptr(self.frames_dev), _stream_ptr()))
            return
        nat.check(self.lib.bpm_gather_frames(_ptr(pcm), nat.PCM_DTYPES[self.np_dtype], self.channels,
                                             _ptr(self.src_items_dev), _host_ptr(self.src_items), self.n_items,
                                             self.src_stride, _ptr(self.frames_dev), _stream_ptr()))

    def ingest(self, staged: torch.Tensor) -> None:
        """One async H2D copy of a pinned staging buffer: the packed kept frames (``pregathered="host"``)
        or the whole batch."""
        dst = self.frames_dev if self.pregathered else self.pcm_dev
        if staged.numel() != dst.numel() or staged.dtype != dst.dtype:
            raise ValueError("staging buffer does not match the runner's frame buffer")
        dst.copy_(staged, non_blocking=True)
        self._pcm_src = dst

    # -- compute
    def launch(self) -> None:
        """Enqueue a1..a4 on the current stream (no host synchronisation)."""
        rc = self.lib.bpm_stage_a(_ptr(self._pcm_src), _ptr(self.items_dev), _host_ptr(self.items), self.n_items,
                                  _ptr(self.design_dev), _host_ptr(self.design_host), self.design_words,
                                  C.byref(self.cfg),
                                  C.byref(self.outs_struct), _ptr(self.ws), self.ws_bytes, _stream_ptr())
        nat.check(rc)

    def result(self) -> StageAResult:
        return StageAResult(self.items, self.plan.rate, self.out)


def _torch_dtype(dt: np.dtype):
    return {np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8,
            np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(dt)]


# ---------------------------------------------------------------------------- single ops
class Ops:
    """Thin per-operator wrappers (one recording or a batch) used by the drop-in functions."""

    def __init__(self):
        self.device = require_cuda()
        self.lib = nat.load_library()

    def _ws(self, nbytes: int) -> torch.Tensor:
        return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)

    def _items(self, lengths: Sequence[int]):
        items = make_items(lengths, lengths)
        dev = torch.from_numpy(items.view(np.int64).reshape(-1, 4).copy()).to(self.device)
        return items, dev

    def frontend(self, pcm: np.ndarray, sample_rate: int, params: Dict, want_debug: bool):
        """K0+K1+K2(+K2b) for one recording -> (envelope, rate, filtered, debug_i16 or None)."""
        plan = plan_filter(sample_rate, params)
        channels = 1 if pcm.ndim == 1 else int(pcm.shape[1])
        if pcm.dtype not in nat.PCM_DTYPES:
            pcm = pcm.astype(np.float64)
        n_in = int(pcm.shape[0])
        if plan.n_dec(n_in) <= PADLEN:
            raise ValueError("The length of the input vector x must be greater than padlen, which is 15.")
        m = plan.m(n_in)
        stride = plan.stride
        pitch_bytes = stride * channels * pcm.dtype.itemsize
        if plan.block == 1 and stride >= 40 and pitch_bytes >= SPARSE_INGEST_MIN_PITCH_BYTES \
                and os.environ.get("BPM_SPARSE_INGEST", "1") != "0":
            # decimate-then-filter touches one frame in `stride`: move only those (bpm_copy_frames, a
            # strided 2-D copy straight out of the caller's array) instead of the whole recording;
            # stage A then reads the compact frames with stride 1 -- bit-identical
            src = np.ascontiguousarray(pcm)
            src_items = make_items([n_in], [m])
            pcm_dev = torch.empty(m * channels, dtype=_torch_dtype(pcm.dtype), device=self.device)
            nat.check(self.lib.bpm_copy_frames(_host_ptr(src), nat.PCM_DTYPES[pcm.dtype], channels, _host_ptr(src_items),
                                               1, stride, _ptr(pcm_dev), _stream_ptr()))
            items, stride = make_items([m], [m]), 1
        else:
            items = make_items([n_in], [m])
            pcm_dev = to_device(pcm.reshape(-1))
        items_dev = torch.from_numpy(items.view(np.int64).reshape(-1, 4).copy()).to(self.device)
        design, design_host = design_images(plan)
        f64 = dict(dtype=torch.float64, device=self.device)
        filt, env, amax = torch.empty(m, **f64), torch.empty(m, **f64), torch.empty(1, **f64)
        nb = int(self.lib.bpm_frontend_workspace_bytes(m, 1))
        ws = self._ws(nb)
        nat.check(self.lib.bpm_frontend(_ptr(pcm_dev), nat.PCM_DTYPES[pcm.dtype], channels, _ptr(items_dev),
                                        _host_ptr(items), 1, stride, _ptr(design), _host_ptr(design_host),
                                        int(design.numel()),
                                        plan.rate // 10, _ptr(filt), _ptr(env), _ptr(amax), _ptr(ws), nb,
                                        _stream_ptr()))
        dbg = None
        if want_debug:
            dbg_dev = torch.empty(m, dtype=torch.int16, device=self.device)
            nat.check(self.lib.bpm_debug_wav(_ptr(filt), _ptr(amax), _ptr(items_dev), _host_ptr(items), 1,
                                             _ptr(dbg_dev), _stream_ptr()))
            dbg = dbg_dev.cpu().numpy()
        return env.cpu().numpy(), plan.rate, filt.cpu().numpy(), dbg

    def quantile(self, x: np.ndarray, q: float) -> float:
        items, items_dev = self._items([len(x)])
        xd = to_device(np.asarray(x, dtype=np.float64))
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        nb = int(self.lib.bpm_quantile_workspace_bytes(1))
        ws = self._ws(nb)
        nat.check(self.lib.bpm_quantile(_ptr(xd), _ptr(items_dev), _host_ptr(items), 1, float(q), _ptr(out),
                                        _ptr(ws), nb, _stream_ptr()))
        return float(out.cpu()[0])

    def find_peaks(self, x: np.ndarray, height: Optional[np.ndarray] = None, prominence: Optional[float] = None,
                   distance: int = 1, sign: int = 1) -> np.ndarray:
        n = len(x)
        items, items_dev = self._items([n])
        xd = to_device(np.asarray(x, dtype=np.float64))
        hd = to_device(np.asarray(height, dtype=np.float64)) if height is not None else None
        pd_ = to_device(np.array([prominence], dtype=np.float64)) if prominence is not None else None
        idx = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        cnt = torch.empty(1, dtype=torch.int64, device=self.device)
        nb = int(self.lib.bpm_find_peaks_workspace_bytes(n, 1))
        ws = self._ws(nb)
        nat.check(self.lib.bpm_find_peaks(_ptr(xd), int(sign), _ptr(hd), _ptr(pd_), int(distance), _ptr(items_dev),
                                          _host_ptr(items), 1, _ptr(idx), _ptr(cnt), _ptr(ws), nb, _stream_ptr()))
        c = int(cnt.cpu()[0])
        return idx[:c].cpu().numpy()

    def rolling_floor(self, env: np.ndarray, knots: np.ndarray, window: int, q: float) -> np.ndarray:
        n = len(env)
        items, items_dev = self._items([n])
        ed = to_device(np.asarray(env, dtype=np.float64))
        kbuf = torch.zeros(max(n, 1), dtype=torch.int64, device=self.device)
        kbuf[:len(knots)] = to_device(np.asarray(knots, dtype=np.int64))
        kc = to_device(np.array([len(knots)], dtype=np.int64))
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        nb = int(self.lib.bpm_rolling_floor_workspace_bytes(n, 1))
        ws = self._ws(nb)
        nat.check(self.lib.bpm_rolling_floor(_ptr(ed), _ptr(kbuf), _ptr(kc), _ptr(items_dev), _host_ptr(items), 1,
                                             int(window), float(q), _ptr(out), _ptr(ws), nb, _stream_ptr()))
        return out.cpu().numpy()

    def noise_floor(self, env: np.ndarray, rate: int, params: Dict):
        n = len(env)
        window = int(params["noise_window_sec"] * rate)
        if window < 3:
            raise ValueError(f"min_periods 3 must be <= window {window}")
        distance = int(params["min_peak_distance_sec"] * rate)
        if distance < 1:
            raise ValueError("`distance` must be greater or equal to 1")
        items, items_dev = self._items([n])
        ed = to_device(np.asarray(env, dtype=np.float64))
        floor = torch.empty(n, dtype=torch.float64, device=self.device)
        tr = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        cnt = torch.empty(1, dtype=torch.int64, device=self.device)
        nb = int(self.lib.bpm_noise_floor_workspace_bytes(n, 1))
        ws = self._ws(nb)
        nat.check(self.lib.bpm_noise_floor(_ptr(ed), _ptr(items_dev), _host_ptr(items), 1, distance,
                                           float(params["trough_prominence_quantile"]),
                                           float(params["noise_floor_quantile"]), window,
                                           float(params.get("trough_rejection_multiplier", 4.0)), _ptr(floor),
                                           _ptr(tr), _ptr(cnt), None, None, _ptr(ws), nb, _stream_ptr()))
        c = int(cnt.cpu()[0])
        return floor.cpu().numpy(), tr[:c].cpu().numpy()

    def raw_peaks_and_metrics(self, env: np.ndarray, floor: np.ndarray, rate: int, params: Dict,
                              with_metrics: bool = True):
        n = len(env)
        distance = int(params["min_peak_distance_sec"] * rate)
        if distance < 1:
            raise ValueError("`distance` must be greater or equal to 1")
        items, items_dev = self._items([n])
        ed = to_device(np.asarray(env, dtype=np.float64))
        fd = to_device(np.asarray(floor, dtype=np.float64))
        pk = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        cnt = torch.empty(1, dtype=torch.int64, device=self.device)
        nb = int(self.lib.bpm_raw_peaks_workspace_bytes(n, 1))
        ws = self._ws(nb)
        nat.check(self.lib.bpm_raw_peaks(_ptr(ed), _ptr(fd), _ptr(items_dev), _host_ptr(items), 1, distance,
                                         float(params["peak_prominence_quantile"]), _ptr(pk), _ptr(cnt), _ptr(ws),
                                         nb, _stream_ptr()))
        if not with_metrics:
            c = int(cnt.cpu()[0])
            return pk[:c].cpu().numpy(), None
        f64 = dict(dtype=torch.float64, device=self.device)
        st, dv, sm = torch.empty(n, **f64), torch.empty(n, **f64), torch.empty(n, **f64)
        nat.check(self.lib.bpm_peak_metrics(_ptr(ed), _ptr(fd), _ptr(pk), _ptr(cnt), _ptr(items_dev),
                                            _host_ptr(items), 1, float(params["deviation_smoothing_factor"]),
                                            _ptr(st), _ptr(dv), _ptr(sm), _stream_ptr()))
        c = int(cnt.cpu()[0])
        d = max(c - 1, 0)
        return pk[:c].cpu().numpy(), {"strength": st[:c].cpu().numpy(), "deviation": dv[:d].cpu().numpy(),
                                      "smoothed": sm[:d].cpu().numpy()}

    def peak_trough_noise(self, env: np.ndarray, floor: np.ndarray, peaks: np.ndarray, troughs: np.ndarray,
                          noise_mult: float, veto_mult: float) -> Dict[str, np.ndarray]:
        n = len(env)
        items, items_dev = self._items([n])
        ed = to_device(np.asarray(env, dtype=np.float64))
        fd = to_device(np.asarray(floor, dtype=np.float64))
        pk = torch.zeros(max(n, 1), dtype=torch.int64, device=self.device)
        tr = torch.zeros(max(n, 1), dtype=torch.int64, device=self.device)
        if len(peaks):
            pk[:len(peaks)] = to_device(np.asarray(peaks, dtype=np.int64))
        if len(troughs):
            tr[:len(troughs)] = to_device(np.asarray(troughs, dtype=np.int64))
        pc = to_device(np.array([len(peaks)], dtype=np.int64))
        tc = to_device(np.array([len(troughs)], dtype=np.int64))
        f64 = dict(dtype=torch.float64, device=self.device)
        pa, na, ra = torch.empty(n, **f64), torch.empty(n, **f64), torch.empty(n, **f64)
        fl = torch.empty(n, dtype=torch.uint8, device=self.device)
        nat.check(self.lib.bpm_peak_trough_noise(_ptr(ed), _ptr(fd), _ptr(pk), _ptr(pc), _ptr(tr), _ptr(tc),
                                                 _ptr(items_dev), _host_ptr(items), 1, float(noise_mult),
                                                 float(veto_mult), _ptr(pa), _ptr(na), _ptr(ra), _ptr(fl),
                                                 _stream_ptr()))
        c = len(peaks)
        return {"prev_amp": pa[:c].cpu().numpy(), "next_amp": na[:c].cpu().numpy(), "ratio": ra[:c].cpu().numpy(),
                "flags": fl[:c].cpu().numpy()}

    # ---- beat-list reductions
    def bpm_series(self, beats: np.ndarray, rate: int, window_us: int):
        b = len(beats)
        items, items_dev = self._items([b])
        bd = to_device(np.asarray(beats, dtype=np.int64))
        f64 = dict(dtype=torch.float64, device=self.device)
        inst, sm, ts = torch.empty(b, **f64), torch.empty(b, **f64), torch.empty(b, **f64)
        us = torch.empty(b, dtype=torch.int64, device=self.device)
        nv = torch.empty(1, dtype=torch.int64, device=self.device)
        nat.check(self.lib.bpm_bpm_series(_ptr(bd), _ptr(items_dev), _host_ptr(items), 1, int(rate), int(window_us),
                                          _ptr(inst), _ptr(sm), _ptr(ts), _ptr(us), _ptr(nv), _stream_ptr()))
        n = int(nv.cpu()[0])
        return inst[:n].cpu().numpy(), sm[:n].cpu().numpy(), ts[:n].cpu().numpy(), us[:n].cpu().numpy()

    def steepest(self, values: np.ndarray, stamps_us: np.ndarray, sign: int, window_sec: float):
        n = len(values)
        items, items_dev = self._items([n])
        vd = to_device(np.asarray(values, dtype=np.float64))
        ud = to_device(np.asarray(stamps_us, dtype=np.int64))
        nv = to_device(np.array([n], dtype=np.int64))
        res = torch.empty(4, dtype=torch.float64, device=self.device)
        ws = self._ws(256)
        nat.check(self.lib.bpm_steepest_slope(_ptr(vd), _ptr(ud), _ptr(nv), _ptr(items_dev), _host_ptr(items), 1,
                                              int(sign), float(window_sec), _ptr(res), _ptr(ws), 256, _stream_ptr()))
        r = res.cpu().numpy()
        if r[0] == 0.0:
            return None
        return int(r[1]), int(r[2]), float(r[3])

    def windowed_hrv(self, beats: np.ndarray, rate: int, win: int, step: int) -> np.ndarray:
        b = len(beats)
        items, items_dev = self._items([b])
        bd = to_device(np.asarray(beats, dtype=np.int64))
        out = torch.empty((max(b, 1), 4), dtype=torch.float64, device=self.device)
        rows = torch.empty(1, dtype=torch.int64, device=self.device)
        nat.check(self.lib.bpm_windowed_hrv(_ptr(bd), _ptr(items_dev), _host_ptr(items), 1, int(rate), int(win),
                                            int(step), _ptr(out), _ptr(rows), _stream_ptr()))
        r = int(rows.cpu()[0])
        return out[:r].cpu().numpy()


def smoothing_window_us(params: Dict) -> int:
    """rolling(f"{output_smoothing_window_sec}s") in microseconds (bpm_analysis.py:1477-1479)."""
    import pandas as pd
    return int(pd.Timedelta(f"{params['output_smoothing_window_sec']}s") // pd.Timedelta(microseconds=1))


def hr_extrema_distance(beats: np.ndarray, rate: int, min_duration_sec: float = 10.0) -> int:
    """`distance` the reference derives for find_major_hr_* from the series index
    (bpm_analysis.py:1492-1494): int((min_duration / 2) / mean gap of the microsecond stamps),
    5 when there is no gap.  Same arithmetic on the beat list the series is made from
    (descriptor-sized host work: one value per beat)."""
    t = np.asarray(beats, dtype=np.int64) / rate
    dt = np.diff(t)
    t1 = t[1:][dt > 1e-6]
    ip = np.trunc(t1)
    us = ip.astype(np.int64) * 1000000 + np.rint((t1 - ip) * 1e6).astype(np.int64)
    gaps = np.diff(us) / 1e6
    if len(gaps) == 0:
        return 5
    mean_gap = np.sum(np.concatenate([[0.0], gaps])) / len(gaps)
    if mean_gap == 0:
        return 5
    return int((min_duration_sec / 2) / mean_gap)


class _BeatSlot:
    """Buffers (and, from the second use on, the CUDA graph) of the a5..a8 chain for ONE shape: number of
    beats, length of the valid series, extrema distance and the window parameters.  The reference calls
    the chain with a handful of list lengths per file (preliminary and final beats); a slot turns the
    ~12 launches, 4 allocations and 3 copies of a call into one graph replay."""

    def __init__(self, owner, b: int, n_series: int, dist: int, rate: int, window_us: int, win: int, step: int):
        self.key = (b, n_series, dist, rate, window_us, win, step)
        self.b, self.n_series, self.dist = b, n_series, dist
        self.rate, self.window_us, self.win, self.step = rate, window_us, win, step
        self.items, self.sitems = make_items([b], [b]), make_items([n_series], [n_series])
        dev = owner.device
        # int64 block on both sides: [list descriptor 4 | series descriptor 4 | beats b | stamps b | tops b | bottoms b | scalars 4]
        self.stage = torch.empty(8 + b, dtype=torch.int64, pin_memory=True)
        sn = self.stage.numpy()
        sn[0:4] = self.items.view(np.int64).reshape(-1)
        sn[4:8] = self.sitems.view(np.int64).reshape(-1)
        self.iblk = torch.empty(8 + 4 * b + 4, dtype=torch.int64, device=dev)
        self.fblk = torch.empty(7 * b + 8 + 1, dtype=torch.float64, device=dev)   # [inst | smoothed | times | hrv 4b | slopes 8 | prominence]
        self.ws_bytes = max(int(owner.lib.bpm_find_peaks_workspace_bytes(b, 1)), 256)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.fh = torch.empty(7 * b + 8, dtype=torch.float64, pin_memory=True)
        self.ih = torch.empty(3 * b + 4, dtype=torch.int64, pin_memory=True)
        self.uses = 0
        self.graph = None

    def enqueue(self, lib, st: int) -> None:
        b, items, sitems = self.b, self.items, self.sitems
        self.iblk[:8 + b].copy_(self.stage, non_blocking=True)
        ip, fp = self.iblk.data_ptr(), self.fblk.data_ptr()
        P = lambda base, off: C.c_void_p(base + 8 * off)                      # noqa: E731 - element offset -> pointer
        items_dev, sitems_dev, bd = P(ip, 0), P(ip, 4), P(ip, 8)
        stamps, tops, bottoms, scal = P(ip, 8 + b), P(ip, 8 + 2 * b), P(ip, 8 + 3 * b), 8 + 4 * b
        inst, smooth, times, hrv, slopes, prom = P(fp, 0), P(fp, b), P(fp, 2 * b), P(fp, 3 * b), P(fp, 7 * b), P(fp, 7 * b + 8)
        nat.check(lib.bpm_bpm_series(bd, items_dev, _host_ptr(items), 1, self.rate, self.window_us, inst, smooth, times,
                                     stamps, P(ip, scal), st))
        # the series holds the valid intervals only (dt > 1e-6 s): its length was counted on the host
        # with the same arithmetic, so the follow-up kernels get an exact descriptor
        nat.check(lib.bpm_steepest_slope(smooth, stamps, P(ip, scal), sitems_dev, _host_ptr(sitems), 1, 0, 20.0, slopes,
                                         _ptr(self.ws), self.ws_bytes, st))
        if self.dist >= 1:
            self.fblk[7 * b + 8:].fill_(5.0)
            for sign, idx, k in ((+1, tops, 1), (-1, bottoms, 2)):
                nat.check(lib.bpm_find_peaks(smooth, sign, None, prom, self.dist, sitems_dev, _host_ptr(sitems), 1, idx,
                                             P(ip, scal + k), _ptr(self.ws), self.ws_bytes, st))
        if b >= self.win:
            nat.check(lib.bpm_windowed_hrv(bd, items_dev, _host_ptr(items), 1, self.rate, self.win, self.step, hrv,
                                           P(ip, scal + 3), st))
        self.fh.copy_(self.fblk[:7 * b + 8], non_blocking=True)
        self.ih.copy_(self.iblk[8 + b:], non_blocking=True)

    def launch(self, owner) -> None:
        stream = owner.stream
        self.uses += 1
        if self.graph is None and self.uses >= 2 and os.environ.get("BPM_DROPIN_GRAPH", "1") != "0":
            try:
                stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                    self.enqueue(owner.lib, torch.cuda.current_stream().cuda_stream)
                self.graph = g
            except Exception:                                    # noqa: BLE001 - the eager path is always valid
                self.graph = False
                torch.cuda.synchronize()
        if self.graph:
            with torch.cuda.stream(stream):
                self.graph.replay()
        else:
            with torch.cuda.stream(stream):
                self.enqueue(owner.lib, stream.cuda_stream)


MAX_BEAT_SLOTS = 8


def beat_chain(owner, beats: np.ndarray, rate: int, window_us: int, win: int, step: int) -> Dict[str, object]:
    """a5..a8 of one beat list in one device round trip (see dropin.DropIn.beat_metrics): ONE upload
    (beat list + the two list descriptors in one pinned block), the kernels, TWO read-backs (one
    float64 block, one int64 block), one synchronisation -- replayed as one CUDA graph from the second
    call with the same shape on."""
    stream = owner.stream
    from .dropin import _Trace
    tr = _Trace(owner.trace)
    b = int(beats.size)
    res: Dict[str, object] = {"n_beats": b}
    if b < 2:
        return res
    dist = hr_extrema_distance(beats, rate)
    tsec = beats / rate
    n_series = int(np.count_nonzero(np.diff(tsec) > 1e-6))          # == n_valid the device will report (:1468)
    res["n_series"] = n_series
    if n_series == 0:
        return res
    slots = owner.__dict__.setdefault("beat_slots", OrderedDict())
    key = (b, n_series, dist, int(rate), int(window_us), int(win), int(step))
    slot = slots.get(key)
    if slot is None:
        slot = slots[key] = _BeatSlot(owner, *key)
        while len(slots) > MAX_BEAT_SLOTS:
            slots.popitem(last=False)
    else:
        slots.move_to_end(key)
    slot.stage.numpy()[8:] = beats
    tr.mark("beat.host_prep")
    slot.launch(owner)
    tr.mark("beat.enqueue")
    stream.synchronize()
    tr.mark("beat.wait")
    f, i = slot.fh.numpy(), slot.ih.numpy()
    nv = int(i[3 * b])
    if nv != n_series:
        raise RuntimeError(f"beat series length: device {nv}, host {n_series}")
    res["hr_distance"] = dist
    res.update(n_valid=nv, inst=f[:nv].copy(), smoothed=f[b:b + nv].copy(), times=f[2 * b:2 * b + nv].copy(),
               stamps=i[:nv].copy(), slopes=f[7 * b:7 * b + 8].copy())
    if dist >= 1:
        res["tops"] = i[b:b + int(i[3 * b + 1])].copy()
        res["bottoms"] = i[2 * b:2 * b + int(i[3 * b + 2])].copy()
    if b >= win:
        res["hrv"] = f[3 * b:3 * b + 4 * int(i[3 * b + 3])].reshape(-1, 4).copy()
    tr.mark("beat.unpack")
    return res


_ops_singleton: Optional[Ops] = None


def ops() -> Ops:
    global _ops_singleton
    if _ops_singleton is None:
        _ops_singleton = Ops()
    return _ops_singleton


class BeatRunner:
    """a5..a8 for one beat list with preallocated device buffers (bpm_bpm_series,
    bpm_steepest_slope x2, bpm_find_peaks x2 for the incline/decline extrema, bpm_windowed_hrv).

    The beat list must be strictly increasing (every interval valid), which is what the
    classifier and the synthetic generators produce; the series length is then B-1.
    """

    def __init__(self, n_beats: int, rate: int, params: Dict, hr_distance: int):
        self.device = require_cuda()
        self.lib = nat.load_library()
        self.rate, self.B = int(rate), int(n_beats)
        if self.B < 3:
            raise ValueError("need at least 3 beats")
        self.window_us = int(round(float(params["output_smoothing_window_sec"]) * 1e6))
        self.win, self.step = int(params["hrv_window_size_beats"]), int(params["hrv_step_size_beats"])
        self.hr_distance = int(hr_distance)
        self.items = make_items([self.B], [self.B])
        self.items_dev = torch.from_numpy(self.items.view(np.int64).reshape(-1, 4).copy()).to(self.device)
        n = self.B - 1
        self.series_items = make_items([n], [n])
        self.series_items_dev = torch.from_numpy(self.series_items.view(np.int64).reshape(-1, 4).copy()).to(self.device)
        f64 = dict(dtype=torch.float64, device=self.device)
        i64 = dict(dtype=torch.int64, device=self.device)
        self.beats_dev = torch.arange(self.B, **i64)         # a valid (strictly increasing) list until upload()
        self.out = {"inst": torch.empty(self.B, **f64), "smoothed": torch.empty(self.B, **f64),
                    "times": torch.empty(self.B, **f64), "stamps": torch.empty(self.B, **i64),
                    "n_valid": torch.empty(1, **i64), "slopes": torch.empty(8, **f64),
                    "tops": torch.empty(self.B, **i64), "n_tops": torch.empty(1, **i64),
                    "bottoms": torch.empty(self.B, **i64), "n_bottoms": torch.empty(1, **i64),
                    "hrv": torch.empty((self.B, 4), **f64), "hrv_rows": torch.empty(1, **i64)}
        self.prom = torch.full((1,), 5.0, **f64)
        self.ws_bytes = int(self.lib.bpm_find_peaks_workspace_bytes(self.B, 1))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)

    def upload(self, beats_host: torch.Tensor) -> None:
        self.beats_dev.copy_(beats_host, non_blocking=True)

    def launch(self) -> None:
        L, o, st = self.lib, self.out, _stream_ptr()
        nat.check(L.bpm_bpm_series(_ptr(self.beats_dev), _ptr(self.items_dev), _host_ptr(self.items), 1, self.rate,
                                   self.window_us, _ptr(o["inst"]), _ptr(o["smoothed"]), _ptr(o["times"]),
                                   _ptr(o["stamps"]), _ptr(o["n_valid"]), st))
        # slopes[0:4] = exertion (+1), slopes[4:8] = recovery (-1): one launch for both directions
        nat.check(L.bpm_steepest_slope(_ptr(o["smoothed"]), _ptr(o["stamps"]), _ptr(o["n_valid"]),
                                       _ptr(self.series_items_dev), _host_ptr(self.series_items), 1, 0, 20.0,
                                       _ptr(o["slopes"]), _ptr(self.ws), self.ws_bytes, st))
        for sign, idx, cnt in ((+1, o["tops"], o["n_tops"]), (-1, o["bottoms"], o["n_bottoms"])):
            nat.check(L.bpm_find_peaks(_ptr(o["smoothed"]), sign, None, _ptr(self.prom), self.hr_distance,
                                       _ptr(self.series_items_dev), _host_ptr(self.series_items), 1, _ptr(idx),
                                       _ptr(cnt), _ptr(self.ws), self.ws_bytes, st))
        nat.check(L.bpm_windowed_hrv(_ptr(self.beats_dev), _ptr(self.items_dev), _host_ptr(self.items), 1, self.rate,
                                     self.win, self.step, _ptr(o["hrv"]), _ptr(o["hrv_rows"]), st))


class GraphedStep:
    """Captures runner launches into one CUDA graph and replays it.

    The runners' buffers are allocated once and the library never allocates or synchronises,
    so a whole step is capturable; replay removes the per-launch CPU / driver overhead of the
    ~60 small kernels of a step.  Runners that do not depend on each other (a1..a4 on the audio,
    a5..a8 on the beat list) are captured on forked streams, so their latency-bound kernels
    overlap on the device.
    """

    def __init__(self, *runners, concurrent: bool = True):
        require_cuda()
        self.runners = runners
        main = torch.cuda.Stream()
        main.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(main):                      # warm-up outside capture (lazy module load etc.)
            for r in runners:
                r.launch()
        torch.cuda.current_stream().wait_stream(main)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        sides = [torch.cuda.Stream() for _ in runners[1:]] if concurrent else []
        with torch.cuda.graph(self.graph):
            cur = torch.cuda.current_stream()
            if concurrent:
                for side, r in zip(sides, runners[1:]):
                    side.wait_stream(cur)                  # fork
                    with torch.cuda.stream(side):
                        r.launch()
                runners[0].launch()
                for side in sides:
                    cur.wait_stream(side)                  # join
            else:
                for r in runners:
                    r.launch()

    def launch(self) -> None:
        self.graph.replay()


def time_pipeline_ingests(n_in: int, sample_rate: int, params: Dict, pcm_pinned: torch.Tensor,
                          beats_pinned: Optional[torch.Tensor] = None, beat_runner_args=None, depth: int = 2,
                          candidates: Sequence[str] = ("host", "ce", "sm"), rounds: int = 6, between=None) -> Dict[str, float]:
    """ms per recording of ``StageAPipeline`` for each ingest, measured on the caller's own recording.
    Which ingest wins depends on the HOST: packing the kept frames needs cores (a box with 16 cores per
    GPU packs a 60-min recording in 0.75 ms, one with 4 cores per GPU in 3 ms), the copy-engine and
    zero-copy forms need PCIe read requests instead (~0.7 G/s per GPU, ~1.9 G/s per root complex).  A
    service measures once at start-up and keeps the fastest.  ``between``: called between candidates
    (a barrier when several ranks measure at the same time, so that they contend as they will later)."""
    out: Dict[str, float] = {}
    for mode in candidates:
        if between is not None:
            between()
        pipe = StageAPipeline(n_in, sample_rate, params, depth=depth, beat_runner_args=beat_runner_args, ingest=mode)

        def run(n):
            for k in range(n):
                if k >= depth:
                    pipe.wait(k - depth)
                pipe.submit(k, pcm_pinned, beats_pinned)
            for k in range(max(0, n - depth), n):
                pipe.wait(k)

        run(depth + 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(rounds)
        out[mode] = (time.perf_counter() - t0) * 1e3 / rounds
        del pipe
    return out


class StageAPipeline:
    """Software pipeline over a stream of equal-shape recordings: ingest | compute | read-back.

    Each of ``depth`` slots owns a pregathered ``StageARunner`` (+ optionally a ``BeatRunner``),
    pinned host result buffers and three events.  ``submit(k, pcm_pinned)`` enqueues, without
    blocking the host,
      ingest stream : the kept frames of the pinned recording cross PCIe -- ``ingest="ce"``: one
                      strided 2-D copy on the copy engine (bpm_copy_frames, default);
                      ``"sm"``: bpm_gather_frames reading mapped pinned memory (zero-copy),
      compute stream: a1..a4 (and a5..a8) as one CUDA-graph replay, after the slot's ingest,
      copy stream   : D2H of every result into the slot's pinned buffers, after the compute;
    ``wait(k)`` blocks until recording k's results are on the host.  With depth >= 2 the PCIe
    reads of recording k+1 overlap the kernels of k and the read-back of k-1.
    """

    LISTS = ("troughs", "peaks", "strength", "deviation", "smoothed_dev")

    def __init__(self, n_in: int, sample_rate: int, params: Dict, depth: int = 2, beat_runner_args=None,
                 pcm_dtype=np.int16, channels: int = 1, use_graph: bool = True, ingest: str = "ce"):
        require_cuda()
        if ingest not in ("ce", "sm", "host"):
            raise ValueError("ingest must be 'ce', 'sm' or 'host'")
        self.ingest = ingest
        self.depth = int(depth)
        self.s_in, self.s_cmp, self.s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        self.slots = []
        for _ in range(self.depth):
            A = StageARunner([n_in], sample_rate, params, pcm_dtype, channels, pregathered=ingest,
                             want_filtered=False)
            Bn = BeatRunner(*beat_runner_args) if beat_runner_args is not None else None
            cap = min(A.total_m, A.total_m // max(int(A.cfg.distance), 1) + 2)   # find_peaks distance bounds the list lengths
            host = {}
            for k, v in A.out.items():
                n = cap if k in self.LISTS else v.numel()
                host[k] = torch.empty(n, dtype=v.dtype).pin_memory()
            stage = (torch.empty(A.total_m * channels, dtype=_torch_dtype(np.dtype(pcm_dtype))).pin_memory()
                     if ingest == "host" else None)
            hostb = {k: torch.empty_like(v, device="cpu").pin_memory() for k, v in Bn.out.items()} if Bn else {}
            with torch.cuda.stream(self.s_cmp):
                A.frames_dev.zero_()
            runners = (A,) if Bn is None else (A, Bn)
            self.slots.append({"A": A, "B": Bn, "cap": cap, "host": host, "hostb": hostb, "runners": runners,
                               "stage": stage,
                               "graph": None, "ev_in": torch.cuda.Event(), "ev_cmp": torch.cuda.Event(),
                               "ev_out": torch.cuda.Event(), "busy": False})
        if ingest == "host":
            from . import classifier
            self.host_lib = classifier.load_host_library()
            self.frame_bytes = int(channels) * np.dtype(pcm_dtype).itemsize
        self.use_graph = use_graph
        self.M = self.slots[0]["A"].total_m
        self.rate = self.slots[0]["A"].plan.rate

    def _graph(self, slot):
        if slot["graph"] is None and self.use_graph:
            torch.cuda.synchronize()
            slot["graph"] = GraphedStep(*slot["runners"])
        return slot["graph"]

    def submit(self, k: int, pcm_pinned: torch.Tensor, beats_pinned: Optional[torch.Tensor] = None) -> None:
        slot = self.slots[k % self.depth]
        if slot["busy"]:
            raise RuntimeError("slot still in flight: call wait() for recording k - depth first")
        A, Bn = slot["A"], slot["B"]
        g = self._graph(slot) if self.use_graph else None
        if self.ingest == "host":
            # K0 by the host cores into this slot's pinned staging buffer (its previous H2D copy finished
            # before the slot's last compute, which wait() has seen complete)
            rc = self.host_lib.bpm_host_gather_frames(C.c_void_p(pcm_pinned.data_ptr()), self.frame_bytes,
                                                      A.total_in, A.src_stride, C.c_void_p(slot["stage"].data_ptr()), 0)
            if rc != 0:
                raise RuntimeError(f"bpm_host_gather_frames failed ({rc})")
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(slot["ev_cmp"])          # the previous compute on this slot has consumed its frames
            if self.ingest == "host":
                A.ingest(slot["stage"])
            else:
                A.gather(pcm_pinned)
            if Bn is not None and beats_pinned is not None:
                Bn.upload(beats_pinned)
            slot["ev_in"].record(self.s_in)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(slot["ev_in"])
            self.s_cmp.wait_event(slot["ev_out"])         # the previous read-back of this slot's outputs is done
            if g is not None:
                g.launch()
            else:
                for r in slot["runners"]:
                    r.launch()
            slot["ev_cmp"].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["ev_cmp"])
            for name, h in slot["host"].items():
                h.copy_(A.out[name][:h.numel()], non_blocking=True)
            for name, h in slot["hostb"].items():
                h.copy_(Bn.out[name], non_blocking=True)
            slot["ev_out"].record(self.s_out)
        slot["busy"] = True

    def wait(self, k: int) -> Dict[str, torch.Tensor]:
        slot = self.slots[k % self.depth]
        slot["ev_out"].synchronize()
        slot["busy"] = False
        out = dict(slot["host"])
        out.update({"beat_" + n: h for n, h in slot["hostb"].items()})
        return out

    def d2h_bytes(self) -> int:
        s = self.slots[0]
        return int(sum(h.numel() * h.element_size() for h in s["host"].values()) +
                   sum(h.numel() * h.element_size() for h in s["hostb"].values()))


def profile_kernels(fn, stream_ptr: Optional[int] = None) -> Dict[str, tuple]:
    """Run ``fn()`` with per-kernel event timing on; returns {kernel: (launches, total_ms)}."""
    lib = nat.load_library()
    nat.check(lib.bpm_profile_begin(stream_ptr if stream_ptr is not None else _stream_ptr()))
    try:
        fn()
    finally:
        buf = C.create_string_buffer(1 << 16)
        nat.check(lib.bpm_profile_end(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out
