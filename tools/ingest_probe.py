#!/usr/bin/env python
"""How fast can the kept frames x[::ds] of a recording in PINNED HOST memory reach the device?

    python tools/ingest_probe.py [duration_sec] [sample_rate]

Compares, for the C2 shape (one int16 frame in 159 is kept), CUDA-event times of
  (a) bpm_gather_frames reading mapped pinned memory from the SMs (what bench.py's e2e uses),
  (b) the copy engine doing the strided copy (cudaMemcpy2DAsync: width 2 B, source pitch 2*ds B),
  (c) a + b splitting the frames between them on two streams,
  (d) one plain copy of the whole recording (the non-sparse alternative),
optionally after cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32).
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpm_analysis_b200 import runtime                     # noqa: E402
from bpm_analysis_b200.params import default_params       # noqa: E402


def timed(fn, streams, reps=5):
    best = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for s in streams[1:]:
            s.wait_event(e0)
        fn()
        for s in streams[1:]:
            ev = torch.cuda.Event()
            ev.record(s)
            streams[0].wait_event(ev)
        e1.record(streams[0])
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    return float(np.median(best))


def main():
    dur = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
    sr = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
    torch.cuda.set_device(0)
    params = default_params()
    n = int(dur * sr)
    runner = runtime.StageARunner([n], sr, params, pregathered=True)
    ds, m = runner.src_stride, runner.total_m
    pinned = torch.empty(n, dtype=torch.int16).pin_memory()
    pinned.numpy()[:] = (np.arange(n, dtype=np.int64) % 30011 - 15000).astype(np.int16)
    cudart = ctypes.CDLL("libcudart.so.12")
    cudart.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    cudart.cudaMemcpy2DAsync.restype = ctypes.c_int
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {"shape": {"n": n, "ds": ds, "m": m}}

    def gather():
        with torch.cuda.stream(s0):
            runner.gather(pinned)
    out["a_sm_gather_ms"] = timed(gather, [s0])
    ref = runner.frames_dev.clone()

    dst = torch.empty(m, dtype=torch.int16, device="cuda")

    def ce(rows, row0=0, stream=s1, width=2):
        rc = cudart.cudaMemcpy2DAsync(dst.data_ptr() + 2 * row0, width, pinned.data_ptr() + 2 * ds * row0, 2 * ds, width,
                                      rows, 1, stream.cuda_stream)
        assert rc == 0, rc
    for rows in (10000, 100000):                              # extrapolate before committing to all rows
        out[f"b_ce_{rows}_rows_ms"] = timed(lambda: ce(rows), [s1], reps=3)
    per_row_us = out["b_ce_100000_rows_ms"] * 1e3 / 100000
    out["b_ce_us_per_row"] = per_row_us
    if per_row_us * m < 50e3:                                  # < 50 ms for the whole recording
        out["b_ce_all_rows_ms"] = timed(lambda: ce(m), [s1], reps=3)
        torch.cuda.synchronize()
        out["b_ce_matches_gather"] = bool(torch.equal(dst.to(torch.float64), ref))
        for frac in (0.25, 0.5):
            k = int(m * frac)
            # SM kernel cannot take a sub-range here, so (c) is bounded below by max(a * (1 - frac), b * frac):
            out[f"c_split_{frac}_bound_ms"] = max(out["a_sm_gather_ms"] * (1 - frac), out["b_ce_all_rows_ms"] * frac)

        def both():
            gather()
            ce(m)
        out["c_both_concurrent_full_ms"] = timed(both, [s0, s1], reps=3)   # contention check: a and b at once
    full = torch.empty(n, dtype=torch.int16, device="cuda")

    def plain():
        with torch.cuda.stream(s0):
            full.copy_(pinned, non_blocking=True)
    out["d_full_copy_ms"] = timed(plain, [s0], reps=3)
    # ---- does the read-back of a step's results (D2H) share a resource with the sparse ingest?
    res_dev = torch.randn(2 * m + 40000, dtype=torch.float64, device="cuda")        # envelope + floor + lists
    res_pin = torch.empty(res_dev.numel(), dtype=torch.float64).pin_memory()
    s2 = torch.cuda.Stream()

    def d2h_ce():
        with torch.cuda.stream(s2):
            res_pin.copy_(res_dev, non_blocking=True)

    class _Alias:                                              # the pinned buffer as a device-visible tensor (UVA)
        __cuda_array_interface__ = {"shape": (res_pin.numel(),), "typestr": "<f8", "data": (res_pin.data_ptr(), False),
                                    "version": 2}
    alias = torch.as_tensor(_Alias(), device="cuda")

    def d2h_sm():                                              # an elementwise kernel storing into mapped host memory
        with torch.cuda.stream(s2):
            torch.add(res_dev, 0.0, out=alias)
    out["e_d2h_bytes"] = int(res_dev.numel() * 8)
    out["e_d2h_ce_alone_ms"] = timed(d2h_ce, [s2])
    out["e_d2h_sm_alone_ms"] = timed(d2h_sm, [s2])
    torch.cuda.synchronize()
    out["e_d2h_sm_correct"] = bool(torch.equal(res_pin, res_dev.cpu()))
    if "b_ce_all_rows_ms" in out:
        out["f_ce_ingest_with_ce_d2h_ms"] = timed(lambda: (ce(m), d2h_ce()), [s1, s2])
        out["f_ce_ingest_with_sm_d2h_ms"] = timed(lambda: (ce(m), d2h_sm()), [s1, s2])
    out["f_sm_ingest_with_ce_d2h_ms"] = timed(lambda: (gather(), d2h_ce()), [s0, s2])
    out["f_sm_ingest_with_sm_d2h_ms"] = timed(lambda: (gather(), d2h_sm()), [s0, s2])
    rc = cudart.cudaDeviceSetLimit(0x05, ctypes.c_size_t(32))
    out["set_l2_fetch_32_rc"] = int(rc)
    out["a_sm_gather_l2fetch32_ms"] = timed(gather, [s0])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
