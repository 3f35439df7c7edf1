"""One long recording as halo-overlapped time chunks, one chunk per GPU.

BASELINE configs[1] / [3]: a 60-min or 24-h stream is split into ``world`` contiguous time
chunks.  Every stage of the front end has bounded support except two global scalars, so a
rank needs its chunk plus a halo, and the ranks exchange three small things:

  stage                       sharded?   needs from outside the chunk            exchange
  --------------------------- ---------- --------------------------------------- ---------------------
  a1 band-pass + envelope      yes        filter transient: ``halo`` kept samples  all_gather(envelope)
     (bpm_analysis.py:1031-1054)          each side (rho^halo < 1e-22) + w/2
  np.quantile (:1067, :225)    replicated whole envelope (8.7 MB at C2)            --
  find_peaks(-env) (:1070)     replicated prominence walks are unbounded           --
  draft floor (:1081-1086)     yes        the troughs around the chunk            --
  sanitisation (:1090-1097)    yes        --                                      all_gather(kept troughs)
  final floor (:1103-1106)     yes        the kept troughs around the chunk       all_gather(floor)
  find_peaks(env, height=floor) replicated --                                      --
  peak metrics (:93-100)       replicated window = f(total peak count)            --

Exactness of the sharded rolling quantile: the floor of a chunk is computed on the sub-range
[a0, a1) of the stream that starts AT a trough at least ``left + 1`` samples before the chunk
and ends AT a trough at least ``off + 1`` samples after it (or at the stream's own ends).
Inside such a range the interpolated trough series equals the global one sample for sample
(np.interp between the same knots), and every window of the chunk's outputs lies inside it,
so the chunk's floor is bit-identical to the unchunked computation, including the reference's
NaN / bfill / ffill behaviour at the true ends of the stream.

The numeric work is done by an *engine* (``DeviceEngine``: libbpm_b200 on CUDA tensors) and
the exchange by a *communicator* (``DistComm``: torch.distributed, NCCL on GPUs;
``ThreadComm``: ranks as threads of one process sharing one GPU, used by the tests).  The CPU
tests drive the same planner with the oracle as the engine under gloo, world size 2.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .dist import shard_range

MIN_TROUGHS = 5          # bpm_analysis.py:1073
MIN_KEPT = 2             # bpm_analysis.py:1102  (len > 2)
FALLBACK_Q = 0.1         # bpm_analysis.py:1114


# --------------------------------------------------------------------------- planning
def filter_halo(spectral_radius: float, block: int, env_window: int, tol: float = 1e-22) -> int:
    """Kept samples after which the transient of a chunk edge has decayed below ``tol``.

    ``spectral_radius`` is per filter input sample and ``block`` the number of filter steps
    per kept sample.  The forward-backward pass convolves two decays (k * rho^k), hence the
    extra log term; the envelope's centred mean adds half its window.
    """
    rho = float(spectral_radius) ** int(block)
    if not (0.0 < rho < 1.0):
        return env_window + 16
    k = math.log(tol) / math.log(rho)
    k += math.log(max(k, 2.0)) / -math.log(rho)
    return int(math.ceil(k)) + env_window + 16


@dataclass(frozen=True)
class ChunkPlan:
    """Chunk geometry in envelope samples (kept samples) and PCM frames."""
    n_frames: int        # N
    m: int               # M = kept samples of the whole stream
    frames_per_sample: int   # stride * block = ds
    halo: int
    world: int

    def core(self, rank: int) -> Tuple[int, int]:
        return shard_range(self.m, self.world, rank)

    def ext(self, rank: int) -> Tuple[int, int]:
        c0, c1 = self.core(rank)
        return max(0, c0 - self.halo), min(self.m, c1 + self.halo)

    def frames(self, rank: int) -> Tuple[int, int]:
        """PCM frames rank ``rank`` has to supply: [f0, f1)."""
        e0, e1 = self.ext(rank)
        return e0 * self.frames_per_sample, min(self.n_frames, e1 * self.frames_per_sample)

    def core_sizes(self) -> List[int]:
        return [self.core(r)[1] - self.core(r)[0] for r in range(self.world)]


def floor_item_range(knots: np.ndarray, c0: int, c1: int, window: int, m: int) -> Tuple[int, int]:
    """Sub-range [a0, a1) of the stream on which the rolling quantile of outputs [c0, c1) is
    bit-identical to the global computation (see module docstring).  ``knots`` ascending."""
    off = (window - 1) // 2
    left = window - 1 - off
    a0, a1 = 0, m
    if len(knots):
        i = int(np.searchsorted(knots, c0 - left - 1, side="right")) - 1     # last knot <= c0-left-1
        if i >= 0:
            a0 = int(knots[i])
        j = int(np.searchsorted(knots, c1 - 1 + off + 1, side="left"))        # first knot >= c1+off
        if j < len(knots):
            a1 = int(knots[j]) + 1
    return a0, a1


# --------------------------------------------------------------------------- communicators
class DistComm:
    """torch.distributed (NCCL on CUDA tensors, gloo on CPU tensors)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def all_gather(self, t: torch.Tensor, sizes: Optional[Sequence[int]] = None) -> List[torch.Tensor]:
        """Ragged all-gather of 1-D tensors: padded to the longest, sliced back by ``sizes``
        (exchanged first when the caller does not know them)."""
        if self.world == 1:
            return [t]
        dist = self._dist
        if sizes is None:
            n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
            ns = [torch.zeros_like(n) for _ in range(self.world)]
            dist.all_gather(ns, n, group=self.group)
            sizes = [int(x.item()) for x in ns]
        cap = max(max(sizes), 1)
        buf = torch.zeros(cap, dtype=t.dtype, device=t.device)
        buf[:t.numel()] = t
        out = torch.empty(self.world * cap, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, buf, group=self.group)
        return [out[r * cap:r * cap + s] for r, s in enumerate(sizes)]


    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            self._dist.all_reduce(t, group=self.group)
        return t

    def all_gather_rows(self, t: torch.Tensor) -> torch.Tensor:
        """Equal-size rows, one per rank -> (world, len) tensor."""
        if self.world == 1:
            return t[None]
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        self._dist.all_gather_into_tensor(out, t.contiguous().view(-1), group=self.group)
        return out.view((self.world,) + tuple(t.shape))


class ThreadComm:
    """``world`` ranks as threads of one process (all on the current device): the exchange is
    a barrier plus a shared slot list.  For tests and for trying a chunking on one GPU."""

    class _Shared:
        def __init__(self, world: int):
            self.barrier = threading.Barrier(world)
            self.slots: List[Optional[torch.Tensor]] = [None] * world

    def __init__(self, shared: "ThreadComm._Shared", rank: int):
        self.shared, self.rank, self.world = shared, rank, len(shared.slots)

    @classmethod
    def make_world(cls, world: int) -> List["ThreadComm"]:
        sh = cls._Shared(world)
        return [cls(sh, r) for r in range(world)]

    def all_gather(self, t: torch.Tensor, sizes: Optional[Sequence[int]] = None) -> List[torch.Tensor]:
        if t.is_cuda:
            torch.cuda.current_stream().synchronize()
        self.shared.slots[self.rank] = t
        self.shared.barrier.wait()
        out = [x.clone() for x in self.shared.slots]
        self.shared.barrier.wait()
        return out


def _thread_all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
    t.copy_(torch.stack(self.all_gather(t)).sum(0))
    return t


def _thread_all_gather_rows(self, t: torch.Tensor) -> torch.Tensor:
    return torch.stack(self.all_gather(t))


ThreadComm.all_reduce_sum = _thread_all_reduce_sum
ThreadComm.all_gather_rows = _thread_all_gather_rows


def run_thread_world(world: int, fn):
    """Run ``fn(comm)`` on ``world`` threads; returns the per-rank results (re-raises errors)."""
    comms = ThreadComm.make_world(world)
    res: List[object] = [None] * world
    errs: List[BaseException] = []
    dev = torch.cuda.current_device() if torch.cuda.is_available() else None

    def body(r):
        try:
            if dev is not None:
                torch.cuda.set_device(dev)
            res[r] = fn(comms[r])
        except BaseException as e:          # noqa: BLE001 - reported to the caller below
            errs.append(e)
            comms[r].shared.barrier.abort()

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    return res


# --------------------------------------------------------------------------- engine (CUDA)
class DeviceEngine:
    """The per-chunk numeric steps as libbpm_b200 calls on CUDA tensors (no numpy round trips)."""

    def __init__(self):
        from . import _native as nat
        from . import runtime
        self.nat, self.rt = nat, runtime
        self.device = runtime.require_cuda()
        self.lib = nat.load_library()
        self._item_cache: Dict[Tuple[int, int], tuple] = {}
        self._ring = torch.empty((self.RING_SLOTS, 16), dtype=torch.int64).pin_memory()
        self._ring_next = 0
        self._ring_lock = threading.Lock()
        # constants of CAPTURED steps must keep their pinned source for the life of the graph: slots of
        # this pool are handed out once and never reused
        self._const_pool = torch.empty((self.CONST_SLOTS, 16), dtype=torch.int64).pin_memory()
        self._const_next = 0
        self.capturing = False

    RING_SLOTS = 256
    CONST_SLOTS = 256

    def small_i64(self, values: Sequence[int]) -> torch.Tensor:
        """A few int64 constants -> device WITHOUT stalling the stream: torch.tensor(list, device=cuda)
        copies from pageable memory, which makes the host wait for the stream.  The values are staged
        in a ring of pinned slots (a slot is reused RING_SLOTS calls later; every step of the chunked
        front end waits for the device at least once, long before that)."""
        with self._ring_lock:
            if self.capturing:
                if self._const_next >= self.CONST_SLOTS:
                    raise RuntimeError("DeviceEngine: out of constant slots for captured steps")
                slot = self._const_pool[self._const_next][:len(values)]
                self._const_next += 1
            else:
                slot = self._ring[self._ring_next][:len(values)]
                self._ring_next = (self._ring_next + 1) % self.RING_SLOTS
        slot.copy_(torch.tensor(list(values), dtype=torch.int64))
        return slot.to(self.device, non_blocking=True)

    # helpers
    def _ws(self, nbytes: int) -> torch.Tensor:
        return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)

    def _items(self, n_in: int, m: int):
        """(host, device) descriptor of one recording; cached (read-only, a handful of shapes per run)"""
        got = self._item_cache.get((n_in, m))
        if got is None:
            items = self.rt.make_items([n_in], [m])
            dev = torch.from_numpy(items.view(np.int64).reshape(-1, 4).copy()).to(self.device)
            if len(self._item_cache) > 256:
                self._item_cache.clear()
            got = self._item_cache[(n_in, m)] = (items, dev)
        return got

    def tensor(self, a: np.ndarray) -> torch.Tensor:
        return self.rt.to_device(np.asarray(a))

    def full(self, n: int, value) -> torch.Tensor:
        return torch.full((n,), float(value), dtype=torch.float64, device=self.device)

    # a1 on a slice of the stream
    def frontend(self, pcm: torch.Tensor, n_in: int, plan, channels: int, np_dtype,
                 want_filtered: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        rt, nat, L = self.rt, self.nat, self.lib
        m = plan.m(n_in)
        items, items_dev = self._items(n_in, m)
        design, design_host = rt.design_images(plan)
        f64 = dict(dtype=torch.float64, device=self.device)
        env, amax = torch.empty(m, **f64), torch.empty(1, **f64)
        # the band-passed signal itself is optional when the envelope is fused into the backward pass
        fused = plan.block == 1 and plan.rate // 10 <= 65
        filt = torch.empty(m, **f64) if (want_filtered or not fused) else None
        nb = int(L.bpm_frontend_workspace_bytes(m, 1))
        ws = self._ws(nb)
        nat.check(L.bpm_frontend(rt._ptr(pcm), nat.PCM_DTYPES[np.dtype(np_dtype)], channels, rt._ptr(items_dev),
                                 rt._host_ptr(items), 1, plan.stride, rt._ptr(design), rt._host_ptr(design_host),
                                 int(design.numel()),
                                 plan.rate // 10, rt._ptr(filt), rt._ptr(env), rt._ptr(amax), rt._ptr(ws), nb,
                                 rt._stream_ptr()))
        self._keep = (ws, items_dev)
        return filt, env

    def quantile(self, x: torch.Tensor, q: float) -> torch.Tensor:
        rt, L = self.rt, self.lib
        items, items_dev = self._items(x.numel(), x.numel())
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        nb = int(L.bpm_quantile_workspace_bytes(1))
        ws = self._ws(nb)
        self.nat.check(L.bpm_quantile(rt._ptr(x), rt._ptr(items_dev), rt._host_ptr(items), 1, float(q), rt._ptr(out),
                                      rt._ptr(ws), nb, rt._stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return out

    def find_peaks(self, x: torch.Tensor, sign: int, height: Optional[torch.Tensor],
                   prominence: Optional[torch.Tensor], distance: int) -> torch.Tensor:
        rt, L = self.rt, self.lib
        n = x.numel()
        items, items_dev = self._items(n, n)
        idx = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        cnt = torch.empty(1, dtype=torch.int64, device=self.device)
        nb = int(L.bpm_find_peaks_workspace_bytes(n, 1))
        ws = self._ws(nb)
        self.nat.check(L.bpm_find_peaks(rt._ptr(x), int(sign), rt._ptr(height), rt._ptr(prominence), int(distance),
                                        rt._ptr(items_dev), rt._host_ptr(items), 1, rt._ptr(idx), rt._ptr(cnt),
                                        rt._ptr(ws), nb, rt._stream_ptr()))
        return idx[:int(cnt.cpu()[0])].clone()

    def rolling_floor(self, env: torch.Tensor, knots: torch.Tensor, window: int, q: float) -> torch.Tensor:
        rt, L = self.rt, self.lib
        n = env.numel()
        items, items_dev = self._items(n, n)
        kbuf = torch.zeros(max(n, 1), dtype=torch.int64, device=self.device)
        kbuf[:knots.numel()] = knots
        kc = torch.tensor([knots.numel()], dtype=torch.int64, device=self.device)
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        nb = int(L.bpm_rolling_floor_workspace_bytes(n, 1))
        ws = self._ws(nb)
        self.nat.check(L.bpm_rolling_floor(rt._ptr(env), rt._ptr(kbuf), rt._ptr(kc), rt._ptr(items_dev),
                                           rt._host_ptr(items), 1, int(window), float(q), rt._ptr(out), rt._ptr(ws),
                                           nb, rt._stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return out

    def sanitize(self, env: torch.Tensor, draft: torch.Tensor, troughs: torch.Tensor, mult: float) -> torch.Tensor:
        rt, L = self.rt, self.lib
        n = env.numel()
        items, items_dev = self._items(n, n)
        tb = torch.zeros(max(n, 1), dtype=torch.int64, device=self.device)
        tb[:troughs.numel()] = troughs
        tc = torch.tensor([troughs.numel()], dtype=torch.int64, device=self.device)
        kept = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        kc = torch.empty(1, dtype=torch.int64, device=self.device)
        nb = int(L.bpm_sanitize_troughs_workspace_bytes(n, 1))
        ws = self._ws(nb)
        self.nat.check(L.bpm_sanitize_troughs(rt._ptr(env), rt._ptr(draft), rt._ptr(tb), rt._ptr(tc),
                                              rt._ptr(items_dev), rt._host_ptr(items), 1, float(mult), rt._ptr(kept),
                                              rt._ptr(kc), rt._ptr(ws), nb, rt._stream_ptr()))
        return kept[:int(kc.cpu()[0])].clone()

    def noise_floor(self, env: torch.Tensor, distance: int, window: int, params: Dict):
        """a2 whole (bpm_noise_floor) on a device envelope -> (floor, kept troughs)."""
        rt, L = self.rt, self.lib
        n = env.numel()
        items, items_dev = self._items(n, n)
        floor = torch.empty(n, dtype=torch.float64, device=self.device)
        tr = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        cnt = torch.empty(1, dtype=torch.int64, device=self.device)
        nb = int(L.bpm_noise_floor_workspace_bytes(n, 1))
        ws = self._ws(nb)
        self.nat.check(L.bpm_noise_floor(rt._ptr(env), rt._ptr(items_dev), rt._host_ptr(items), 1, int(distance),
                                         float(params["trough_prominence_quantile"]),
                                         float(params["noise_floor_quantile"]), int(window),
                                         float(params.get("trough_rejection_multiplier", 4.0)), rt._ptr(floor),
                                         rt._ptr(tr), rt._ptr(cnt), None, None, rt._ptr(ws), nb, rt._stream_ptr()))
        return floor, tr[:int(cnt.cpu()[0])].clone()

    def raw_peaks(self, env: torch.Tensor, floor: torch.Tensor, distance: int, prom_q: float) -> torch.Tensor:
        """a3 (bpm_raw_peaks) on device tensors."""
        rt, L = self.rt, self.lib
        n = env.numel()
        items, items_dev = self._items(n, n)
        pk = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        cnt = torch.empty(1, dtype=torch.int64, device=self.device)
        nb = int(L.bpm_raw_peaks_workspace_bytes(n, 1))
        ws = self._ws(nb)
        self.nat.check(L.bpm_raw_peaks(rt._ptr(env), rt._ptr(floor), rt._ptr(items_dev), rt._host_ptr(items), 1,
                                       int(distance), float(prom_q), rt._ptr(pk), rt._ptr(cnt), rt._ptr(ws), nb,
                                       rt._stream_ptr()))
        return pk[:int(cnt.cpu()[0])].clone()

    def peak_metrics(self, env: torch.Tensor, floor: torch.Tensor, peaks: torch.Tensor, factor: float):
        rt, L = self.rt, self.lib
        n = env.numel()
        items, items_dev = self._items(n, n)
        pk = torch.zeros(max(n, 1), dtype=torch.int64, device=self.device)
        pk[:peaks.numel()] = peaks
        cnt = torch.tensor([peaks.numel()], dtype=torch.int64, device=self.device)
        f64 = dict(dtype=torch.float64, device=self.device)
        st, dv, sm = torch.empty(n, **f64), torch.empty(n, **f64), torch.empty(n, **f64)
        self.nat.check(L.bpm_peak_metrics(rt._ptr(env), rt._ptr(floor), rt._ptr(pk), rt._ptr(cnt), rt._ptr(items_dev),
                                          rt._host_ptr(items), 1, float(factor), rt._ptr(st), rt._ptr(dv), rt._ptr(sm),
                                          rt._stream_ptr()))
        torch.cuda.current_stream().synchronize()
        c = peaks.numel()
        d = max(c - 1, 0)
        return st[:c], dv[:d], sm[:d]


    # ---- chunk mode (ShardedFrontEnd): nothing here synchronises the host
    def key_state(self, n_total: int, k: int) -> torch.Tensor:
        return self.small_i64([0, int(k), int(n_total)])

    def key_histogram(self, x: torch.Tensor, shift: int, bits: int, state: torch.Tensor, hist: torch.Tensor) -> None:
        self.nat.check(self.lib.bpm_key_histogram(self.rt._ptr(x), x.numel(), int(shift), int(bits), 0,
                                                  self.rt._ptr(state), self.rt._ptr(hist), self.rt._stream_ptr()))

    def key_pick(self, hist: torch.Tensor, bits: int, state: torch.Tensor) -> None:
        self.nat.check(self.lib.bpm_key_pick(self.rt._ptr(hist), int(bits), self.rt._ptr(state), self.rt._stream_ptr()))

    def key_collect(self, x: torch.Tensor, up_shift: int, state: torch.Tensor, cap: int) -> torch.Tensor:
        """-> int64[cap + 4] (bit patterns of uint64): [count, smallest key above the bucket, smallest and
        largest key in it, keys...]"""
        out = torch.zeros(cap + 4, dtype=torch.int64, device=self.device)
        out[1:3] = -1                                            # ~0 as uint64
        base = out.data_ptr()
        self.nat.check(self.lib.bpm_key_collect(self.rt._ptr(x), x.numel(), int(up_shift), 0, self.rt._ptr(state),
                                                int(cap), C.c_void_p(base + 32), C.c_void_p(base),
                                                self.rt._stream_ptr()))
        return out

    def key_finish(self, rows: torch.Tensor, cap: int, state: torch.Tensor, gamma: float, out: torch.Tensor,
                   status: torch.Tensor) -> None:
        self.nat.check(self.lib.bpm_key_finish(self.rt._ptr(rows), int(rows.shape[0]), int(cap), self.rt._ptr(state),
                                               float(gamma), self.rt._ptr(out), self.rt._ptr(status),
                                               self.rt._stream_ptr()))

    def chunk_chain(self, env: torch.Tensor, thr: torch.Tensor, qstat: torch.Tensor, geom: "ChunkGeometry",
                    params: Dict) -> Dict[str, torch.Tensor]:
        """a2 + a3 + strength (:95) of one chunk + halo and the chunk's proof, enqueued back to back.
        thr = (trough, peak) prominence thresholds on the device."""
        rt, L, nat = self.rt, self.lib, self.nat
        n = env.numel()
        items, items_dev = self._items(n, n)
        f64 = dict(dtype=torch.float64, device=self.device)
        i64 = dict(dtype=torch.int64, device=self.device)
        floor, strength = torch.empty(n, **f64), torch.empty(n, **f64)
        kept, every, peaks = torch.empty(n, **i64), torch.empty(n, **i64), torch.empty(n, **i64)
        big = np.iinfo(np.int64).max
        # counts {kept, all, peaks} | flags {edge hits, trough anchors l/r, peak anchors l/r} | proof[8]
        head = self.small_i64([0, 0, 0, 0, -1, big, -1, big, 0, 0, 0, 0, 0, 0, 0, 0])
        hb = head.data_ptr()
        cnt, flg, prf = hb, hb + 24, hb + 64
        nb = max(int(L.bpm_noise_floor_chunk_workspace_bytes(n)), int(L.bpm_find_peaks_workspace_bytes(n, 1)))
        ws = self._ws(nb)
        st = rt._stream_ptr()
        thr_p = thr.data_ptr()
        ends = (int(not geom.at_start), int(not geom.at_end))
        nat.check(L.bpm_noise_floor_chunk(rt._ptr(env), rt._ptr(items_dev), rt._host_ptr(items), int(geom.distance),
                                          C.c_void_p(thr_p), float(params["noise_floor_quantile"]), int(geom.window),
                                          float(params.get("trough_rejection_multiplier", 4.0)),
                                          int(geom.t_lo), int(geom.t_hi), ends[0], ends[1], rt._ptr(floor), rt._ptr(kept),
                                          C.c_void_p(cnt), rt._ptr(every), C.c_void_p(cnt + 8), C.c_void_p(flg),
                                          C.c_void_p(flg + 8), rt._ptr(ws), nb, st))
        nat.check(L.bpm_find_peaks_chunk(rt._ptr(env), 1, rt._ptr(floor), C.c_void_p(thr_p + 8), int(geom.distance),
                                         rt._ptr(items_dev), rt._host_ptr(items), int(geom.core_lo), int(geom.core_hi),
                                         ends[0], ends[1], rt._ptr(peaks), C.c_void_p(cnt + 16), C.c_void_p(flg),
                                         C.c_void_p(flg + 24), rt._ptr(ws), nb, st))
        nat.check(L.bpm_peak_strength(rt._ptr(env), rt._ptr(floor), rt._ptr(peaks), C.c_void_p(cnt + 16),
                                      rt._ptr(items_dev), rt._host_ptr(items), 1, rt._ptr(strength), st))
        nat.check(L.bpm_chunk_proof(rt._ptr(every), rt._ptr(kept), rt._ptr(peaks), C.c_void_p(cnt), C.c_void_p(flg),
                                    rt._ptr(qstat), n, int(geom.core_lo), int(geom.core_hi), int(geom.t_lo),
                                    int(geom.t_hi), int(geom.at_start), int(geom.at_end), int(geom.filter_halo),
                                    int(geom.distance), int(geom.window), C.c_void_p(prf), st))
        self._keep_chunk = (ws, items_dev)
        return {"floor": floor, "kept": kept, "every": every, "peaks": peaks, "strength": strength,
                "head": head, "proof": head[8:16]}

    def chunk_pack(self, c: Dict[str, torch.Tensor], origin: int, cap_t: int, cap_p: int) -> torch.Tensor:
        out = torch.empty(cap_t + 2 * cap_p, dtype=torch.int64, device=self.device)
        self.nat.check(self.lib.bpm_chunk_pack(self.rt._ptr(c["kept"]), self.rt._ptr(c["peaks"]),
                                               self.rt._ptr(c["strength"]), self.rt._ptr(c["proof"]), int(origin),
                                               int(cap_t), int(cap_p), self.rt._ptr(out), self.rt._stream_ptr()))
        return out

    def chunk_unpack(self, rows: torch.Tensor, table: torch.Tensor, cap_t: int, cap_p: int, n_t: int, n_p: int):
        troughs = torch.empty(n_t, dtype=torch.int64, device=self.device)
        peaks = torch.empty(n_p, dtype=torch.int64, device=self.device)
        strength = torch.empty(n_p, dtype=torch.float64, device=self.device)
        self.nat.check(self.lib.bpm_chunk_unpack(self.rt._ptr(rows), self.rt._ptr(table), int(rows.shape[0]), int(cap_t),
                                                 int(cap_p), self.rt._ptr(troughs), self.rt._ptr(peaks),
                                                 self.rt._ptr(strength), self.rt._stream_ptr()))
        return troughs, peaks, strength

    def deviation_series(self, strength: torch.Tensor, factor: float):
        """deviation and its rolling mean (:96-100) of a whole strength list"""
        rt, L = self.rt, self.lib
        c = strength.numel()
        d = max(c - 1, 0)
        if c < 2:
            z = torch.zeros(0, dtype=torch.float64, device=self.device)
            return z, z
        items, items_dev = self._items(c, c)
        cnt = self.small_i64([c])
        f64 = dict(dtype=torch.float64, device=self.device)
        dv, sm = torch.empty(2 * c, **f64), torch.empty(c, **f64)
        self.nat.check(L.bpm_deviation_series(rt._ptr(strength.contiguous()), rt._ptr(cnt), rt._ptr(items_dev),
                                              rt._host_ptr(items), 1, float(factor), rt._ptr(dv), rt._ptr(sm),
                                              rt._stream_ptr()))
        self._keep_dev = (items_dev, cnt)
        return dv[:d], sm[:d]


# --------------------------------------------------------------------------- stream-wide order statistics
KEY_PASSES = ((53, 11), (42, 11), (31, 11), (20, 11), (9, 11), (0, 9))     # (shift, bits), most significant first
DESCENT_PASSES = 3               # digits resolved before the bucket is gathered: sign + exponent + 21 mantissa bits
COLLECT_CAP = 4096               # keys a rank contributes to the final bucket (and the bucket's limit)


def quantile_position(n_total: int, q: float) -> Tuple[int, float]:
    """numpy's virtual index (linear method): element k and the fraction towards element k + 1."""
    v = np.float64(n_total - 1) * np.float64(q)
    k = int(min(max(np.floor(v), 0), n_total - 1))
    return k, float(v - np.floor(v))


def stream_quantiles(engine, comm, x_core: torch.Tensor, n_total: int, qs: Sequence[float]):
    """np.quantile(x, q) (linear method) for each q of a series whose samples are spread over the
    ranks (each rank passes ITS samples, every sample on exactly one rank): a radix descent over
    the order-preserving 64-bit keys with the per-digit histograms summed over the ranks, then
    the few keys left in the bucket are gathered and ordered.  Everything is enqueued -- the
    descent's state lives on the device -- and every rank ends up with the same values.
    -> (values float64[len(qs)], status int64[len(qs)]: nonzero = bucket too large, value invalid)"""
    if n_total < 1:
        raise ValueError("quantile of an empty stream")
    nq = len(qs)
    out = torch.zeros(nq, dtype=torch.float64, device=x_core.device)
    status = torch.ones(nq, dtype=torch.int64, device=x_core.device)
    pos = [quantile_position(n_total, q) for q in qs]
    states = [engine.key_state(n_total, k) for k, _ in pos]
    for p in range(DESCENT_PASSES):
        shift, bits = KEY_PASSES[p]
        hist = torch.zeros((nq, 1 << bits), dtype=torch.int64, device=x_core.device)
        for i in range(nq if p else 1):                          # the first digit's histogram is the same for all
            engine.key_histogram(x_core, shift, bits, states[i], hist[i])
        if p == 0 and nq > 1:
            hist[1:] = hist[0]
        comm.all_reduce_sum(hist)
        for i in range(nq):
            engine.key_pick(hist[i], bits, states[i])
    up = KEY_PASSES[DESCENT_PASSES - 1][0]
    mine = torch.stack([engine.key_collect(x_core, up, states[i], COLLECT_CAP) for i in range(nq)])
    rows = comm.all_gather_rows(mine)                            # (world, nq, cap + 4)
    for i in range(nq):
        engine.key_finish(rows[:, i].contiguous(), COLLECT_CAP, states[i], pos[i][1], out[i:i + 1], status[i:i + 1])
    return out, status


@dataclass(frozen=True)
class ChunkGeometry:
    """One rank's chunk in indices local to its extended range [0, n)."""
    n: int
    core_lo: int
    core_hi: int
    t_lo: int            # the range the trough list is claimed on
    t_hi: int
    at_start: bool
    at_end: bool
    filter_halo: int
    distance: int
    window: int


# --------------------------------------------------------------------------- the chunked front end
class ChunkedFrontEnd:
    """a1..a4 of ONE recording, time-chunked over ``comm.world`` ranks.

    Every rank constructs it with the same arguments, feeds the PCM frames ``frames()`` asks
    for, and receives the full result (every rank ends up with the whole envelope, floor and
    lists, which is what the sequential classifier that follows needs).
    """

    def __init__(self, n_frames: int, sample_rate: int, params: Dict, comm, engine, plan=None,
                 pcm_dtype=np.int16, channels: int = 1):
        if plan is None:
            from .runtime import plan_filter
            plan = plan_filter(sample_rate, params)
        self.plan = plan
        self.np_dtype, self.channels = np.dtype(pcm_dtype), int(channels)
        self._common(plan.rate, params, comm, engine)
        m = plan.m(int(n_frames))
        halo = filter_halo(plan.design.spectral_radius, plan.block, plan.rate // 10)
        self.chunks = ChunkPlan(int(n_frames), m, plan.stride * plan.block, halo, comm.world)
        c0, c1 = self.chunks.core(comm.rank)
        if c1 - c0 < 1:
            raise ValueError("more ranks than envelope samples")

    @classmethod
    def for_envelope(cls, m: int, rate: int, params: Dict, comm, engine) -> "ChunkedFrontEnd":
        """Only the analysis stages (``analyse``) on an envelope every rank already holds."""
        self = cls.__new__(cls)
        self.plan = None
        self._common(int(rate), params, comm, engine)
        self.chunks = ChunkPlan(int(m), int(m), 1, 0, comm.world)
        return self

    def _common(self, rate: int, params: Dict, comm, engine) -> None:
        self.params, self.comm, self.engine, self.rate = params, comm, engine, rate
        self.window = int(params["noise_window_sec"] * rate)                       # :1083
        if self.window < 3:
            raise ValueError(f"min_periods 3 must be <= window {self.window}")
        self.distance = int(params["min_peak_distance_sec"] * rate)               # :226, :1066
        if self.distance < 1:
            raise ValueError("`distance` must be greater or equal to 1")

    def frames(self) -> Tuple[int, int]:
        return self.chunks.frames(self.comm.rank)

    # -- a1
    def envelope(self, pcm_slice) -> Tuple[torch.Tensor, torch.Tensor]:
        """Filter + envelope of this rank's frames -> (full envelope, this rank's filtered core)."""
        ch, r = self.chunks, self.comm.rank
        f0, f1 = ch.frames(r)
        (c0, c1), (e0, _) = ch.core(r), ch.ext(r)
        filt, env = self.engine.frontend(pcm_slice, f1 - f0, self.plan, self.channels, self.np_dtype)
        core_env = env[c0 - e0:c1 - e0].contiguous()
        parts = self.comm.all_gather(core_env, ch.core_sizes())
        return torch.cat(parts), filt[c0 - e0:c1 - e0]

    # -- a2 + a3 + a4 on a full envelope every rank holds
    def _chunk_floor(self, env: torch.Tensor, knots_host: np.ndarray) -> Tuple[torch.Tensor, int, int]:
        c0, c1 = self.chunks.core(self.comm.rank)
        a0, a1 = floor_item_range(knots_host, c0, c1, self.window, self.chunks.m)
        lo, hi = np.searchsorted(knots_host, [a0, a1])
        local = self.engine.tensor(knots_host[lo:hi] - a0)
        item = self.engine.rolling_floor(env[a0:a1], local, self.window, float(self.params["noise_floor_quantile"]))
        return item, a0, a1

    def analyse(self, env: torch.Tensor) -> Dict[str, torch.Tensor]:
        """_calculate_dynamic_noise_floor + _find_raw_peaks + metrics (bpm_analysis.py:1064-1117,
        :223-229, :93-100) with the two rolling quantiles and the sanitisation sharded by chunk."""
        E, P, ch, comm = self.engine, self.params, self.chunks, self.comm
        c0, c1 = ch.core(comm.rank)
        sizes = ch.core_sizes()
        q_tp = E.quantile(env, float(P["trough_prominence_quantile"]))                            # :1067
        all_troughs = E.find_peaks(env, -1, None, q_tp, self.distance)                            # :1070
        if all_troughs.numel() < MIN_TROUGHS:                                                     # :1073-1077
            floor = E.full(ch.m, float(E.quantile(env, float(P["noise_floor_quantile"]))[0]))
            troughs = all_troughs
        else:
            knots = all_troughs.cpu().numpy()
            draft_item, a0, a1 = self._chunk_floor(env, knots)                                    # :1081-1086
            lo, hi = np.searchsorted(knots, [c0, c1])
            mine = E.tensor(knots[lo:hi] - a0)
            kept_local = E.sanitize(env[a0:a1], draft_item, mine,
                                    float(P.get("trough_rejection_multiplier", 4.0)))             # :1090-1097
            troughs = torch.cat(comm.all_gather(kept_local + a0))
            if troughs.numel() > MIN_KEPT:                                                        # :1102-1106
                floor_item, a0, a1 = self._chunk_floor(env, troughs.cpu().numpy())
            else:                                                                                 # :1107-1110
                floor_item = draft_item
            floor = torch.cat(comm.all_gather(floor_item[c0 - a0:c1 - a0].contiguous(), sizes))
            if bool(torch.isnan(floor).all()):                                                    # :1113-1115
                floor = E.full(ch.m, float(E.quantile(env, FALLBACK_Q)[0]))
        q_pp = q_tp if P["peak_prominence_quantile"] == P["trough_prominence_quantile"] else \
            E.quantile(env, float(P["peak_prominence_quantile"]))                                 # :225
        peaks = E.find_peaks(env, +1, floor, q_pp, self.distance)                                 # :227
        strength, deviation, smoothed = E.peak_metrics(env, floor, peaks, float(P["deviation_smoothing_factor"]))
        return {"envelope": env, "floor": floor, "troughs": troughs, "peaks": peaks, "strength": strength,
                "deviation": deviation, "smoothed_dev": smoothed}

    def run(self, pcm_slice) -> Dict[str, torch.Tensor]:
        env, filt_core = self.envelope(pcm_slice)
        out = self.analyse(env)
        out["filtered_core"] = filt_core
        return out


# --------------------------------------------------------------------------- every stage sharded
DISTANCE_MARGIN_HOPS = 32        # samples of margin per unit of `distance` in which an anchor is looked for


class ShardedFrontEnd(ChunkedFrontEnd):
    """a1..a4 of ONE recording with EVERY stage evaluated per time chunk; the ranks exchange digit
    histograms (the stream-wide np.quantile thresholds), a handful of counters, and at the end the
    per-chunk trough / peak / strength lists -- never the envelope (``gather_series=False``).

    A rank evaluates its core [c0, c1) on the extended range [c0 - H, c1 + H), H = filter halo +
    analysis halo, and PROVES afterwards, from what it computed, that nothing it reports can depend on
    samples it did not see:

      find_peaks     the kernels count every core decision that touched an open end (``edge_hits``:
                     prominence walks, flat runs, unresolved distance chains) and report the anchors
                     next to the core (a candidate that outranks everything within ``distance``; no
                     distance-rule dependency crosses one) -- see bpm_b200.h, bpm_find_peaks_chunk
      rolling floors the window of an output lies between two knots (troughs) of the part of the
                     trough list that is itself proven, so the interpolated series under the window
                     is the stream's (np.interp between the same knots)
      count rules    :1073 (< 5 troughs), :1102 (<= 2 kept) are decided on the stream's totals

    If any rank cannot prove its chunk (or a count rule fires) ALL ranks fall back to
    ChunkedFrontEnd (envelope gathered, global steps replicated), which is exact by construction;
    ``result["sharded"]`` says which path produced the result.
    """

    def __init__(self, n_frames: int, sample_rate: int, params: Dict, comm, engine, plan=None,
                 pcm_dtype=np.int16, channels: int = 1, analysis_halo: Optional[int] = None):
        super().__init__(n_frames, sample_rate, params, comm, engine, plan, pcm_dtype, channels)
        self.filter_halo = self.chunks.halo
        self._graph = None
        self._set_halo(analysis_halo)

    @classmethod
    def for_envelope(cls, m: int, rate: int, params: Dict, comm, engine,
                     analysis_halo: Optional[int] = None) -> "ShardedFrontEnd":
        self = super().for_envelope(m, rate, params, comm, engine)
        self.filter_halo = 0
        self._graph = None
        self._set_halo(analysis_halo)
        return self

    def _set_halo(self, analysis_halo: Optional[int]) -> None:
        self.margin = DISTANCE_MARGIN_HOPS * self.distance + 64
        if analysis_halo is None:
            # two rolling windows deep (draft floor -> kept troughs -> final floor), each half a window
            # plus the gap to the next trough (allowed: another half window), and three margins
            analysis_halo = 2 * self.window + 3 * self.margin
        self.analysis_halo = int(analysis_halo)
        ch = self.chunks
        self.chunks = ChunkPlan(ch.n_frames, ch.m, ch.frames_per_sample, self.filter_halo + self.analysis_halo, ch.world)

    def geometry(self) -> ChunkGeometry:
        ch, r = self.chunks, self.comm.rank
        (c0, c1), (e0, e1) = ch.core(r), ch.ext(r)
        n = e1 - e0
        at_start, at_end = e0 == 0, e1 == ch.m
        doubt = self.filter_halo + self.margin                   # zone next to an open end nothing is claimed about
        return ChunkGeometry(n, c0 - e0, c1 - e0, 0 if at_start else doubt, n if at_end else n - doubt, at_start,
                             at_end, self.filter_halo, self.distance, self.window)

    # -- the step in two parts around its one wait for the device
    def _enqueue(self, env_ext: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Everything up to the all-gathered table of proofs, enqueued (no host synchronisation)."""
        E, P, ch, comm = self.engine, self.params, self.chunks, self.comm
        g = self.geometry()
        env_core = env_ext[g.core_lo:g.core_hi]
        # stream-wide thresholds (:1067, :225)
        q_t, q_p = float(P["trough_prominence_quantile"]), float(P["peak_prominence_quantile"])
        if q_t == q_p:
            one, qstat = stream_quantiles(E, comm, env_core, ch.m, [q_t])
            thr, qstat = one.repeat(2), qstat.repeat(2)
        else:
            thr, qstat = stream_quantiles(E, comm, env_core, ch.m, [q_t, q_p])
        c = E.chunk_chain(env_ext, thr, qstat, g, P)
        c.update(env=env_ext, env_core=env_core, thr=thr, table=comm.all_gather_rows(c["proof"]).contiguous())
        return c

    def _finish(self, c: Dict[str, torch.Tensor]) -> Optional[Dict[str, torch.Tensor]]:
        """The wait, the verdict of all ranks, the one list exchange, the stream-wide deviation series."""
        E, P, ch, comm = self.engine, self.params, self.chunks, self.comm
        e0 = ch.ext(comm.rank)[0]
        g = self.geometry()
        table = c["table"].cpu().numpy()                          # the chunk's one wait for the device
        mine = table[comm.rank]
        self.last_proof = {"ok": not bool(mine[0]), "proven_floor": (int(mine[6]) + e0, int(mine[7]) + e0),
                           "table": table}
        if table[:, 0].any() or table[:, 1].sum() < MIN_TROUGHS or table[:, 2].sum() <= MIN_KEPT:
            return None
        # ONE exchange for the three lists: [kept troughs | peaks | strength bits], padded to the longest
        cap_t, cap_p = int(table[:, 2].max()), int(table[:, 3].max())
        rows = comm.all_gather_rows(E.chunk_pack(c, e0, cap_t, cap_p))
        troughs, peaks_all, strength_all = E.chunk_unpack(rows, c["table"], cap_t, cap_p, int(table[:, 2].sum()),
                                                          int(table[:, 3].sum()))
        deviation, smoothed = E.deviation_series(strength_all, float(P["deviation_smoothing_factor"]))
        return {"troughs": troughs, "peaks": peaks_all, "strength": strength_all, "deviation": deviation,
                "smoothed_dev": smoothed, "envelope_core": c["env_core"],
                "floor_core": c["floor"][g.core_lo:g.core_hi], "thresholds": c["thr"]}

    def analyse_sharded(self, env_ext: torch.Tensor) -> Optional[Dict[str, torch.Tensor]]:
        """Per-chunk evaluation; None if some rank could not prove its chunk.  The host waits for the
        device ONCE (for the table of proofs and list lengths of all ranks)."""
        return self._finish(self._enqueue(env_ext))

    def _complete(self, out: Optional[Dict[str, torch.Tensor]], env_ext: torch.Tensor,
                  gather_series: bool) -> Dict[str, torch.Tensor]:
        ch, comm = self.chunks, self.comm
        (c0, c1), (e0, _) = ch.core(comm.rank), ch.ext(comm.rank)
        if out is None:                                           # exact by construction, slower
            env = torch.cat(comm.all_gather(env_ext[c0 - e0:c1 - e0].contiguous(), ch.core_sizes()))
            out = ChunkedFrontEnd.analyse(self, env)
            out["envelope_core"], out["floor_core"] = out["envelope"][c0:c1], out["floor"][c0:c1]
            out["sharded"] = False
            return out
        out["sharded"] = True
        if gather_series:
            sizes = ch.core_sizes()
            out["envelope"] = torch.cat(comm.all_gather(out["envelope_core"].contiguous(), sizes))
            out["floor"] = torch.cat(comm.all_gather(out["floor_core"].contiguous(), sizes))
        return out

    def analyse_local(self, env_ext: torch.Tensor, gather_series: bool = False) -> Dict[str, torch.Tensor]:
        return self._complete(self.analyse_sharded(env_ext), env_ext, gather_series)

    def analyse(self, env: torch.Tensor, gather_series: bool = True) -> Dict[str, torch.Tensor]:
        """On an envelope every rank already holds (for_envelope)."""
        e0, e1 = self.chunks.ext(self.comm.rank)
        return self.analyse_local(env[e0:e1], gather_series)

    def _first_part(self, pcm_slice, want_filtered: bool):
        f0, f1 = self.chunks.frames(self.comm.rank)
        filt, env = self.engine.frontend(pcm_slice, f1 - f0, self.plan, self.channels, self.np_dtype, want_filtered)
        c = self._enqueue(env)
        c["filt"] = filt
        return c

    def _captured(self, pcm_slice, want_filtered: bool):
        """The step's first part as ONE CUDA graph (its NCCL collectives included): a replay costs the
        host ~20 us instead of ~0.7 ms of Python and launches, which is what bounds a step once the
        chunks are short (8 GPUs: 0.75 ms of kernels per rank).  Captured for one input tensor; a
        different tensor (or any capture failure) falls back to the eager path."""
        key = (pcm_slice.data_ptr(), pcm_slice.numel(), bool(want_filtered))
        st = self._graph
        if st is not None and st["key"] == key:
            st["graph"].replay()
            return st["out"]
        if st is not None and st.get("failed"):
            return self._first_part(pcm_slice, want_filtered)
        E = self.engine
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                        # warm-up: lazy initialisations, NCCL set-up
                self._first_part(pcm_slice, want_filtered)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            E.capturing = True
            try:
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    out = self._first_part(pcm_slice, want_filtered)
            finally:
                E.capturing = False
            self._graph = {"key": key, "graph": graph, "out": out}
            graph.replay()
            return out
        except Exception as e:                                   # noqa: BLE001 - the eager path is always valid
            import warnings
            warnings.warn(f"ShardedFrontEnd: CUDA-graph capture of the step failed ({e!r}); running eagerly")
            self._graph = {"key": None, "failed": True}
            torch.cuda.synchronize()
            return self._first_part(pcm_slice, want_filtered)

    def run(self, pcm_slice, gather_series: bool = False, want_filtered: bool = True,
            graph: bool = False) -> Dict[str, torch.Tensor]:
        """graph=True: replay the captured first part (DistComm / one rank only; the ranks of a thread
        world synchronise through the host and cannot be captured)."""
        ch, r = self.chunks, self.comm.rank
        (c0, c1), (e0, _) = ch.core(r), ch.ext(r)
        if graph and isinstance(self.comm, DistComm):
            c = self._captured(pcm_slice, want_filtered)
        else:
            c = self._first_part(pcm_slice, want_filtered)
        out = self._complete(self._finish(c), c["env"], gather_series)
        if c["filt"] is not None:
            out["filtered_core"] = c["filt"][c0 - e0:c1 - e0]
        return out

