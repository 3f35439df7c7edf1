#!/usr/bin/env python
"""Headline benchmark: audio-hours/s of the bpm_analysis front end on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--filter-mode parity|fullrate]

Workload (BASELINE.json configs[1], "C2"): one synthetic 60-min 48 kHz mono int16
heart-sound recording (60 -> 170 -> 80 BPM ramp) per GPU.  One step = one pass of the hot
path over it: a1..a4 (band-pass + envelope, dynamic noise floor, raw peaks, per-peak
metrics) on the audio, then a5..a8 (BPM series, steepest slopes, incline/decline extrema,
windowed HRV) on the recording's beat list.  With N > 1 (torchrun) every rank processes its
own recording of the same shape -- recordings are independent units, so there is no
data-path collective ("weak" scaling); NCCL is only used for the barrier and the
max-over-ranks of the device time.

`value` is measured with the PCM already in HBM; `e2e` through the public runner objects
with pinned HOST buffers (H2D of the PCM and beat list and D2H of every result inside the
timed region).  `--impl reference` times the CPU oracle (oracle/ref_port.py: the reference's
own numpy/scipy/pandas calls) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC, UNIT = "audio_hours_per_sec", "audio-hours/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--filter-mode", default="parity", choices=["parity", "fullrate"])
    ap.add_argument("--duration-sec", type=float, default=3600.0)
    ap.add_argument("--sample-rate", type=int, default=48000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream", action="store_true", help="skip the `stream` sub-record (one 24-h recording "
                    "time-chunked over the GPUs of this run)")
    ap.add_argument("--stream-hours", type=float, default=24.0)
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip the fullrate / sweep / stream sub-records")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--no-zero-copy", action="store_true", help="e2e: copy the whole recording to the device first")
    ap.add_argument("--e2e-depth", type=int, default=3, help="e2e: recordings in flight (1 = strictly serial steps)")
    ap.add_argument("--e2e-ingest", default="auto", choices=["auto", "host", "ce", "sm"],
                    help="e2e_pipelined: the kept frames are packed by the host cores and copied (host), cross PCIe as a "
                         "strided copy-engine copy (ce), or are read from mapped pinned memory by a kernel (sm); auto "
                         "(default) measures the three on this box before the timed region and keeps the fastest")
    ap.add_argument("--dump-kernels", default=None, help="write the per-kernel table to this JSON file")
    ap.add_argument("--workload", default="recordings", choices=["recordings", "stream", "batch", "holter", "sweep"],
                    help="recordings: one C2 recording per GPU (weak scaling, the headline). stream: ONE C2 "
                         "recording split into halo-overlapped time chunks over the GPUs (strong scaling). "
                         "batch: BASELINE configs[2], one rank's share = 128 x 10-min 44.1 kHz recordings in one "
                         "call. holter: configs[3], one 24-h 4 kHz recording (M = 28.8 M envelope samples). sweep: "
                         "configs[4], 256 band-pass / noise-floor settings over one 30-min 48 kHz recording")
    return ap.parse_args()


def workload_name(args) -> str:
    return (f"C2: synthetic {args.duration_sec / 60:g}-min {args.sample_rate / 1000:g} kHz mono int16 heart-sound "
            f"recording (60->170->80 BPM ramp), one per GPU")


def make_recording(args, rank: int):
    from bpm_analysis_b200 import synth
    return synth.config_c2(seed=2 + rank, duration_sec=args.duration_sec, sample_rate=args.sample_rate)


def bench_params(args):
    from bpm_analysis_b200.params import default_params
    p = default_params()
    p["save_filtered_wav"] = False
    p["filter_mode"] = args.filter_mode
    return p


def svc_cap(M: int, A) -> int:
    """entries per list read-back of a drop-in step (find_peaks distance bounds the list lengths)"""
    return min(M, M // max(int(A.cfg.distance), 1) + 2)


def hr_extrema_distance(beat_idx: np.ndarray, rate: int) -> int:
    """`distance` the reference derives for find_major_hr_* (bpm_analysis.py:1492-1494)."""
    t = beat_idx[1:] / rate
    ip = np.trunc(t)
    us = ip.astype(np.int64) * 1000000 + np.rint((t - ip) * 1e6).astype(np.int64)
    gaps = np.diff(us) / 1e6
    mean_gap = np.sum(np.concatenate([[0.0], gaps])) / len(gaps)
    return max(1, int((10 / 2) / mean_gap))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- roofline
def algorithmic_bytes(kernel: str, shp: dict, mode: str = "parity") -> float:
    """Compulsory bytes per launch of one kernel at its own interface:
    every input element it needs read once + every output element written once."""
    N, M, T, P, B = shp["N"], shp["M"], shp["T"], shp["P"], shp["B"]
    C = shp.get("C", 4 * max(T, P))                      # local maxima before the distance / prominence rules
    s_in = 2
    table = {
        "k_contract_i16": N * s_in + 8 * M,            # N*s_in + 8*M
        "k_contract_generic": M * s_in + 8 * M,        # parity: only the kept samples are needed
        # mean of the forward and the backward pass.  parity: fwd M*s_in in, x 8 + s_f 32 out; bwd x 8 + s_f 32 in,
        # y 8 out.  fullrate: fwd uf 32 in, s_f 32 out; bwd ub0 32 + s_f 32 + x 8 in, y 8 out
        "k_scan": (0.5 * ((s_in + 40) + 48) if mode == "parity" else 0.5 * (64 + 80)) * M,
        "k_gather_frames": M * s_in + 8 * M,
        "k_envelope": 16 * M,
        "k_select_pass": 8 * M,
        "k_select_next": 8 * M,
        # parity mode: fwd reads M*s_in frames, writes y_f (8 M); bwd reads y_f, writes the envelope (8 M)
        "k_sos_fwd": (s_in + 8) * M, "k_sos_bwd": 16 * M,
        "k_localmax_compact": 8 * M + 8 * C, "k_distance_tiles": 17 * C, "k_prominence_compact": 17 * C + 8 * max(T, P),
        "k_sanitize_compact": 24 * T,
        "k_knot_table": 40 * T,
        "k_rolling_floor": 16 * T + 8 * M,             # SURVEY §8(d) K5+K6
        "k_rolling_floor_blk": 16 * T + 8 * M,
        "k_find_peaks_small": 16 * B,
        "k_sanitize_flags": 24 * T,
        "k_peak_strength": 32 * P, "k_peak_deviation": 16 * P, "k_dev_smooth_slide": 16 * P,
        "k_select_collect": 8 * M, "k_select_finish": 8 * 4096,
        # the chunked stream's extra kernels (per rank: M = its chunk)
        "k_key_hist": 8 * M, "k_key_collect": 8 * M, "k_chunk_pack": 24 * (T + 2 * P), "k_chunk_unpack": 24 * (T + 2 * P),
        "k_bpm_instant": 32 * B, "k_bpm_smooth": 24 * B, "k_steepest": 16 * B, "k_hrv": 8 * B + 32 * (B // 5),
    }
    return float(table.get(kernel, 0.0))


PIPE_INGEST = {
    "host": "bpm_host_gather_frames: the host cores pack the kept frames x[::ds] into pinned staging, one cudaMemcpyAsync "
            "moves them (h2d bytes = the kept frames); ",
    "ce": "bpm_copy_frames: the kept frames x[::ds] leave pinned host memory as one strided 2-D copy on the copy engine "
          "(h2d bytes = the kept frames); ",
    "sm": "zero-copy: bpm_gather_frames reads the kept frames from pinned host memory (h2d bytes = 32-byte sector per "
          "kept frame); ",
}

BEAT_BRANCH_KERNELS = {"k_steepest", "k_bpm_instant", "k_bpm_smooth", "k_hrv"}
# the three find_peaks kernels are launched four times per step: twice on the envelope (troughs, raw peaks) and
# twice on the ~6 k-sample BPM series of the beat branch (a7's extrema), where a launch is all overhead
BEAT_BRANCH_LAUNCHES = {"k_localmax_compact": 2, "k_distance_tiles": 2, "k_prominence_compact": 2}

ROOFLINE_NOTES = {
    "k_rolling_floor_blk": "bounded by shared-memory latency / instruction issue, not HBM: an exact rolling quantile "
                           "(sample sort + sliding rank pointer per CTA) whose algorithmic traffic is 8 B per output; "
                           "the HBM fraction is reported because the contract asks for it (DESIGN.md section 4)",
    "k_contract_i16": "HBM and FP64-pipe bound together: 8 DFMA per 2-byte sample",
    "k_sos_fwd": "forward cascade scan over the extended signal; gathers the strided PCM itself (ld.global.L2::64B: one "
                 "64-byte DRAM fetch per 2-byte kept frame -- the algorithmic bytes count the 2 bytes)",
    "k_sos_bwd": "backward cascade scan with |y| and the centred rolling mean formed in its epilogue; the band-passed "
                 "signal never reaches HBM",
}


def measured_peak_gbs():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str, mode: str):
    """dram bytes per launch from the committed ncu --set full capture, if one exists."""
    try:
        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh).get(mode, {}).get(kernel)
    except Exception:
        return None


# --------------------------------------------------------------------------- parity of the run's own outputs
def parity_report(got: dict, oracle_out, A) -> dict:
    """The drop-in step's outputs on the bench recording against the CPU oracle on the same
    recording ('parity asserted in the same run').  Float signals: max|d| /
    max|ref| <= 1e-9; index lists and beat-list reductions: exact."""
    fe, br = oracle_out

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        if a.shape != b.shape:
            return float("inf")
        if a.size == 0:
            return 0.0
        return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) or 1.0))

    def same_records(x, y):
        xs, ys = (x, y) if isinstance(x, list) else ([x], [y])
        if (x is None) != (y is None) or len(xs) != len(ys):
            return False
        if x is None:
            return True
        return all(list(p) == list(q) and all((abs(p[k] - q[k]) <= 1e-9 * max(1.0, abs(q[k]))) if isinstance(p[k], float)
                                              else p[k] == q[k] for k in p) for p, q in zip(xs, ys))

    rep = {"tolerance": 1e-9,
           "envelope_rel_err": rel(got["envelope"], fe["envelope"]),
           "floor_rel_err": rel(got["floor"].values, fe["floor"]),
           "troughs_exact": bool(np.array_equal(got["troughs"], fe["troughs"])),
           "raw_peaks_exact": bool(np.array_equal(got["peaks"], fe["peaks"]) and np.array_equal(got["prelim_peaks"], fe["peaks"])),
           "smoothed_deviation_rel_err": rel(got["smoothed_dev"].values, fe["smoothed_dev_series"].values) if "smoothed_dev_series" in fe else None,
           "bpm_series_index_exact": bool(got["smoothed_bpm"].index.equals(br["smoothed_bpm"].index)),
           "bpm_series_rel_err": rel(got["smoothed_bpm"].values, br["smoothed_bpm"].values),
           "slopes_exact": bool(same_records(got["peak_recovery_stats"], br["peak_recovery_stats"]) and
                                same_records(got["peak_exertion_stats"], br["peak_exertion_stats"])),
           "inclines_declines_exact": bool(same_records(got["major_inclines"], br["major_inclines"]) and
                                           same_records(got["major_declines"], br["major_declines"])),
           "hrv_rel_err": rel(got["windowed_hrv_df"].values, br["windowed_hrv_df"].values),
           "n_troughs": int(len(fe["troughs"])), "n_raw_peaks": int(len(fe["peaks"])),
           "n_hrv_rows": int(len(br["windowed_hrv_df"]))}
    floats = [v for k, v in rep.items() if k.endswith("rel_err") and v is not None]
    rep["ok"] = bool(all(v <= 1e-9 for v in floats) and all(v for k, v in rep.items() if k.endswith("exact")))
    return rep


# --------------------------------------------------------------------------- CPU side
def _cpu_one(job, keep: bool = False):
    pcm, sr, beat_idx, params = job
    from oracle import ref_port
    fe = ref_port.front_end(pcm, sr, params)
    br = ref_port.beat_reductions(beat_idx, fe["rate"], params)
    return (fe, br) if keep else len(fe["peaks"])


# ---- the UNMODIFIED reference (baseline/_ref, see baseline/install_ref.sh) on the host cores
_REF_JOB = None


def _reference_one(_):
    """One recording through the reference's own functions for a1..a8 (bpm_analysis.py:1731-1732,
    :1635 -> :85-111, :223-229, :1704-1710), WAV file in, pandas objects out."""
    path, out_dir, beat_idx, params = _REF_JOB
    from baseline import ref_loader
    ref = ref_loader.load()
    env, rate = ref.preprocess_audio(path, params, out_dir)
    floor, troughs = ref._calculate_dynamic_noise_floor(env, rate, params)
    clf = ref.PeakClassifier(env, rate, params, None, floor, troughs, None, None)      # _initialize_state + _find_raw_peaks
    m = {}
    m["smoothed_bpm"], m["bpm_times"] = ref.calculate_bpm_series(beat_idx, rate, params)
    m["major_inclines"] = ref.find_major_hr_inclines(m["smoothed_bpm"])
    m["major_declines"] = ref.find_major_hr_declines(m["smoothed_bpm"])
    m["peak_recovery_stats"] = ref.find_peak_recovery_rate(m["smoothed_bpm"])
    m["peak_exertion_stats"] = ref.find_peak_exertion_rate(m["smoothed_bpm"])
    m["windowed_hrv_df"] = ref.calculate_windowed_hrv(beat_idx, rate, params)
    return len(clf.state["all_peaks"])


_POOL_JOB = None


def _cpu_pool_worker(_):
    return _cpu_one(_POOL_JOB)


def cpu_step(pcm, sr, beat_idx, params, workers: int, keep: bool = False):
    """One CPU step: `workers` recordings, one per process (the reference is single-threaded)."""
    global _POOL_JOB
    if workers <= 1:
        t0 = time.perf_counter()
        res = _cpu_one((pcm, sr, beat_idx, params), keep)
        dt = time.perf_counter() - t0
        return (dt, res) if keep else dt
    import multiprocessing as mp
    _POOL_JOB = (pcm, sr, beat_idx, params)
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_pool_worker, range(workers))            # spin-up excluded
        t0 = time.perf_counter()
        pool.map(_cpu_pool_worker, range(workers))
        return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    params = bench_params(args)
    from bpm_analysis_b200 import synth
    from bpm_analysis_b200.params import effective_decimation
    pcm, sr, beats = make_recording(args, 0)
    _, rate, _ = effective_decimation(sr, params)
    cores = os.cpu_count() or 1
    workers = min(cores, 64 if args.filter_mode == "parity" else 8)
    # bounded sample: keep (steps + warmup) CPU steps inside ~3 minutes (~7 s per audio-hour per core)
    budget = 170.0 / max(1, args.steps + args.warmup + 1)
    sample_sec = float(min(args.duration_sec, max(120.0, budget / 8.0 * 3600.0)))
    n = int(sample_sec * sr)
    pcm_s = pcm[:n]
    beat_idx = synth.beats_to_envelope_indices(beats[beats < sample_sec - 1.0], rate)
    import multiprocessing as mp
    import tempfile
    from baseline import ref_loader
    global _POOL_JOB, _REF_JOB
    use_ref = ref_loader.available() and args.filter_mode == "parity"
    tmp = None
    if use_ref:
        # the reference's entry point reads a WAV file (bpm_analysis.py:1014): the sample is written once,
        # outside the timed region, to shared memory
        from scipy.io import wavfile
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        tmp = tempfile.mkdtemp(prefix="bpm_ref_", dir=base)
        path = os.path.join(tmp, "sample.wav")
        wavfile.write(path, sr, pcm_s)
        ref_params = dict(ref_loader.default_params(), save_filtered_wav=False)
        _REF_JOB = (path, tmp, beat_idx, ref_params)
        worker, kind = _reference_one, "reference"
        what = ("the UNMODIFIED reference (baseline/_ref): preprocess_audio on a WAV in /dev/shm, "
                "_calculate_dynamic_noise_floor, PeakClassifier.__init__ (_initialize_state + _find_raw_peaks), "
                "calculate_bpm_series, find_major_hr_inclines/declines, find_peak_recovery/exertion_rate, "
                "calculate_windowed_hrv")
    else:
        _POOL_JOB = (pcm_s, sr, beat_idx, params)
        worker, kind = _cpu_pool_worker, "port"
        what = "oracle/ref_port.py (the reference's numpy/scipy/pandas calls restated), a1..a8"
    ctx = mp.get_context("fork")
    times = []
    try:
        with ctx.Pool(workers) as pool:
            for _ in range(args.warmup):
                pool.map(worker, range(workers))
            for _ in range(args.steps):
                t0 = time.perf_counter()
                pool.map(worker, range(workers))
                times.append(time.perf_counter() - t0)
    finally:
        if tmp:
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)
    step = sum(times) / len(times)
    value = workers * (sample_sec / 3600.0) / step
    sample = (f"{workers} x first {sample_sec:g} s of the C2 recording per step, one process per recording "
              f"(numpy/scipy/pandas are single-threaded on this path); {what}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "filter_mode": args.filter_mode},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------- GPU side
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    l2_fetch = os.environ.get("BPM_L2_FETCH_GRANULARITY")     # experiment knob (bytes: 32 / 64 / 128), off by default
    if l2_fetch:
        import ctypes
        torch.cuda.init()
        rc = ctypes.CDLL("libcudart.so.12").cudaDeviceSetLimit(0x05, ctypes.c_size_t(int(l2_fetch)))
        sys.stderr.write(f"cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, {l2_fetch}) -> {rc}\n")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from bpm_analysis_b200.dist import bind_to_gpu_numa
        bind_to_gpu_numa(local_rank)                # pinned ingest buffers on the GPU's own socket

    from bpm_analysis_b200 import _native, synth
    from bpm_analysis_b200.runtime import BeatRunner, GraphedStep, StageARunner, profile_kernels

    lib = _native.load_library()
    params = bench_params(args)
    pcm, sr, beats = make_recording(args, rank)
    # the product path (preprocess_audio -> envelope) never reads the band-passed signal itself: in the
    # decimate-first order it does not leave the SMs (want_filtered=False); fullrate mode still writes it
    A = StageARunner([len(pcm)], sr, params, want_filtered=(args.filter_mode != "parity"))
    rate = A.plan.rate
    beat_idx = synth.beats_to_envelope_indices(beats, rate)
    Bn = BeatRunner(len(beat_idx), rate, params, hr_extrema_distance(beat_idx, rate))
    pcm_pin = torch.from_numpy(pcm).pin_memory()
    beats_pin = torch.from_numpy(beat_idx).pin_memory()
    A.upload_pinned(pcm_pin)
    Bn.upload(beats_pin)
    torch.cuda.synchronize()
    audio_hours = args.duration_sec / 3600.0

    def step_eager():
        A.launch()
        Bn.launch()

    graphed = None if args.no_graph else GraphedStep(A, Bn)
    step = step_eager if graphed is None else graphed.launch

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # kernels per step (counted on an eager step; a graph replay launches the same kernels)
    l_before = lib.bpm_launch_count()
    step_eager()
    torch.cuda.synchronize()
    launches_per_step = int(lib.bpm_launch_count() - l_before)

    # ---- device-resident timing
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = lib.bpm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = int(lib.bpm_launch_count() - l0)
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = world * audio_hours / (ms_step / 1e3)

    # ---- end to end: pinned host in, pinned host out, every step
    M = A.total_m
    host = {k: torch.empty_like(A.out[k], device="cpu").pin_memory()
            for k in ("envelope", "floor", "troughs", "peaks", "trough_count", "peak_count", "strength",
                      "smoothed_dev")}
    hostb = {k: torch.empty_like(v, device="cpu").pin_memory() for k, v in Bn.out.items()}

    # decimate-then-filter touches every ds-th frame only: the ingest kernel pulls those frames out
    # of pinned host memory (zero-copy), and a 2-deep pipeline overlaps the PCIe-bound ingest of
    # step k+1 with the kernels of step k and the read-back of step k-1 (runtime.StageAPipeline).
    # Every step still moves its own input host->device and its own results device->host inside
    # the timed region; nothing is reused between steps.
    zero_copy = (args.filter_mode == "parity") and not args.no_zero_copy
    e2e_mode = "pipelined zero-copy" if zero_copy else "serial full copy"
    if zero_copy:
        from bpm_analysis_b200.runtime import StageAPipeline
        depth = max(1, args.e2e_depth)
        beat_args = (len(beat_idx), rate, params, hr_extrema_distance(beat_idx, rate))
        ingest_ms = None
        if args.e2e_ingest == "auto":
            # which ingest is fastest is a property of the host (cores per GPU, PCIe root complexes): measured
            # here, untimed, with all ranks of the run measuring the same candidate at the same time
            from bpm_analysis_b200.runtime import time_pipeline_ingests
            ingest_ms = time_pipeline_ingests(len(pcm), sr, params, pcm_pin, beats_pin, beat_args, depth=depth,
                                              between=barrier)
            ingest_ms = {k: max_over_ranks(v) for k, v in ingest_ms.items()}
            args.e2e_ingest = min(ingest_ms, key=ingest_ms.get)
        pipe = StageAPipeline(len(pcm), sr, params, depth=depth, beat_runner_args=beat_args,
                              use_graph=not args.no_graph, ingest=args.e2e_ingest)

        def e2e_run(n_steps):
            res = None
            for k in range(n_steps):
                if k >= depth:
                    res = pipe.wait(k - depth)
                pipe.submit(k, pcm_pin, beats_pin)
            for k in range(max(0, n_steps - depth), n_steps):
                res = pipe.wait(k)
            return res

        res = e2e_run(max(depth + 1, args.warmup))
        barrier()
        t0 = time.perf_counter()
        res = e2e_run(args.steps)
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
        nt, npk = int(res["trough_count"][0]), int(res["peak_count"][0])
        nv, rows = int(res["beat_n_valid"][0]), int(res["beat_hrv_rows"][0])
        # the pipelined results are the device-resident run's results
        assert nt == int(A.out["trough_count"][0]) and npk == int(A.out["peak_count"][0])
        assert torch.equal(res["peaks"][:npk], A.out["peaks"][:npk].cpu())
        assert torch.equal(res["floor"], A.out["floor"].cpu())
        # sm: one 32-byte sector per kept frame crosses PCIe; ce: the 2-D copy moves the kept frames themselves
        h2d = M * (32 if args.e2e_ingest == "sm" else pcm_pin.element_size()) + beats_pin.numel() * 8
        d2h = pipe.d2h_bytes()
        del pipe
    else:
        def e2e_step():
            A.upload_pinned(pcm_pin)
            Bn.upload(beats_pin)
            step()
            for k in ("trough_count", "peak_count"):
                host[k].copy_(A.out[k], non_blocking=True)
            for k in ("n_tops", "n_bottoms", "hrv_rows", "slopes", "n_valid"):
                hostb[k].copy_(Bn.out[k], non_blocking=True)
            host["envelope"].copy_(A.out["envelope"], non_blocking=True)
            host["floor"].copy_(A.out["floor"], non_blocking=True)
            torch.cuda.synchronize()
            nt, npk = int(host["trough_count"][0]), int(host["peak_count"][0])
            host["troughs"][:nt].copy_(A.out["troughs"][:nt], non_blocking=True)
            for k in ("peaks", "strength", "smoothed_dev"):
                host[k][:npk].copy_(A.out[k][:npk], non_blocking=True)
            nv, rows = int(hostb["n_valid"][0]), int(hostb["hrv_rows"][0])
            for k in ("smoothed", "times", "stamps"):
                hostb[k][:nv].copy_(Bn.out[k][:nv], non_blocking=True)
            hostb["tops"][:int(hostb["n_tops"][0])].copy_(Bn.out["tops"][:int(hostb["n_tops"][0])], non_blocking=True)
            hostb["bottoms"][:int(hostb["n_bottoms"][0])].copy_(Bn.out["bottoms"][:int(hostb["n_bottoms"][0])],
                                                                non_blocking=True)
            hostb["hrv"][:rows].copy_(Bn.out["hrv"][:rows], non_blocking=True)
            torch.cuda.synchronize()
            return nt, npk, nv, rows

        for _ in range(max(1, args.warmup // 2)):
            nt, npk, nv, rows = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
        h2d = pcm_pin.numel() * pcm_pin.element_size() + beats_pin.numel() * 8
        d2h = 2 * M * 8 + nt * 8 + npk * 24 + nv * 24 + rows * 32 + 8 * 8 + 7 * 8 + \
            (int(hostb["n_tops"][0]) + int(hostb["n_bottoms"][0])) * 8
    pipe_ms, pipe_h2d, pipe_d2h = e2e_ms, h2d, d2h

    # ---- end to end through the DROP-IN functions: what a caller of the reference's names gets.
    # One recording at a time, the ten frontend.* calls in analyze_wav_file's order (bpm_analysis.py:
    # 1731-1732, :1635, :1649-1650, :1740, :1704-1710), host arrays in, host arrays / pandas objects
    # out, no pipelining across recordings.  The session layer (dropin.py) is emptied at the top of
    # every step, so nothing is carried from one step to the next; the preliminary pass gets its own
    # (thinned) beat list, as in the reference, so both BPM-series calls do their work.
    from bpm_analysis_b200 import frontend
    from bpm_analysis_b200.dropin import dropin as _dropin
    svc = _dropin()
    pcm_host = pcm_pin.numpy()
    prelim_beats = beat_idx[(np.arange(len(beat_idx)) % 9) != 4]

    class _Clf:                                       # the attributes _initialize_state / _find_raw_peaks read
        pass

    def dropin_step():
        svc.forget()
        clf = _Clf()
        env, r, _, _ = frontend.preprocess_pcm(pcm_host, sr, params, want_filtered=False)
        floor, troughs = frontend._calculate_dynamic_noise_floor(env, r, params)
        clf.audio_envelope, clf.sample_rate, clf.params = env, r, params
        st1 = frontend._initialize_state(clf, None, floor, troughs)
        ps, pt = frontend.calculate_bpm_series(prelim_beats, r, params)
        phase = frontend.find_recovery_phase(ps, pt, params)
        st2 = frontend._initialize_state(clf, 80.0, floor, troughs)
        sm, bt = frontend.calculate_bpm_series(beat_idx, r, params)
        res = {"envelope": env, "floor": floor, "troughs": troughs, "peaks": st2["all_peaks"],
               "smoothed_dev": st2["smoothed_dev_series"], "prelim_peaks": st1["all_peaks"], "phase": phase,
               "smoothed_bpm": sm, "bpm_times": bt,
               "major_inclines": frontend.find_major_hr_inclines(sm),
               "major_declines": frontend.find_major_hr_declines(sm),
               "hrr_stats": frontend.calculate_hrr(sm),
               "peak_recovery_stats": frontend.find_peak_recovery_rate(sm),
               "peak_exertion_stats": frontend.find_peak_exertion_rate(sm),
               "windowed_hrv_df": frontend.calculate_windowed_hrv(beat_idx, r, params)}
        return res

    import logging
    logging.getLogger().setLevel(logging.ERROR)
    for _ in range(max(3, args.warmup)):
        dres = dropin_step()
    barrier()
    s0 = dict(svc.stats)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dres = dropin_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    s1 = dict(svc.stats)
    assert s1["stage_a_calls"] - s0["stage_a_calls"] == args.steps
    nt_d, npk_d = len(dres["troughs"]), len(dres["peaks"])
    nb, nb2 = len(beat_idx), len(prelim_beats)
    h2d = M * pcm_pin.element_size() + (nb + nb2) * 8 + 2 * 64            # kept frames + two beat lists + descriptors
    d2h = (2 * M * 8 + 5 * svc_cap(M, A) * 8 + 4 * 8) + (7 * nb + 8 + 3 * nb + 4) * 8 + (7 * nb2 + 8 + 3 * nb2 + 4) * 8
    clocks = sampler.stop()
    e2e_value = world * audio_hours / (e2e_ms / 1e3)

    # ---- per-kernel event timing (separate pass: events between launches perturb the step)
    roofline, kernels = None, {}
    if not args.no_profile:
        prof = profile_kernels(lambda: [step_eager() for _ in range(args.steps)])
        torch.cuda.synchronize()
        shp = {"N": len(pcm), "M": M, "T": nt, "P": npk, "B": len(beat_idx)}
        peak, peak_src = measured_peak_gbs()
        total_ms = sum(v[1] for v in prof.values()) or 1.0
        for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
            avg_us = ms * 1e3 / cnt
            ab = algorithmic_bytes(name, shp, args.filter_mode)
            kernels[name] = {"launches_per_step": cnt / args.steps, "avg_us": round(avg_us, 3),
                             "share": round(ms / total_ms, 4), "alg_bytes": ab,
                             "gbs": round(ab / (avg_us * 1e-6) / 1e9, 2) if avg_us > 0 else None}
        # the beat-list kernels (a5..a8) run on a forked graph branch beside the audio branch and are off
        # the critical path: the roofline line describes the dominant kernel of the audio branch
        def audio_time(name):
            kk = kernels[name]
            return kk["avg_us"] * max(kk["launches_per_step"] - BEAT_BRANCH_LAUNCHES.get(name, 0), 0)
        top = max((name for name in kernels if name not in BEAT_BRANCH_KERNELS), key=audio_time)
        k = kernels[top]
        roofline = {"kernel": top, "bound": "hbm", "achieved": k["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": round(k["gbs"] / peak, 5) if k["gbs"] else None,
                    "traffic": ncu_traffic(top, args.filter_mode), "peak_source": peak_src,
                    "share_of_step": k["share"], "avg_launch_us": k["avg_us"],
                    "how": "CUDA events after every launch over a separate pass of `steps` steps",
                    "note": ROOFLINE_NOTES.get(top, "")}
        if args.dump_kernels and rank == 0:
            with open(args.dump_kernels, "w") as fh:
                json.dump({"filter_mode": args.filter_mode, "shape": shp, "ms_per_step": ms_step, "kernels": kernels},
                          fh, indent=1)

    # ---- CPU baseline beside it (rank 0, single GPU runs only) + parity of this run's GPU outputs
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        sample_sec = min(args.duration_sec, 3600.0)
        n = int(sample_sec * sr)
        bi = synth.beats_to_envelope_indices(beats[beats < sample_sec - 1.0], rate)
        reps = 4                                     # ~11 s of CPU work at the default size
        t_first, oracle_out = cpu_step(pcm[:n], sr, bi, params, 1, keep=True)
        ts = [t_first] + [cpu_step(pcm[:n], sr, bi, params, 1) for _ in range(reps - 1)]
        t = sum(ts) / reps
        if n == len(pcm):
            parity = parity_report(dres, oracle_out, A)
        cpu = {"value": (sample_sec / 3600.0) / t, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{reps} x first {sample_sec:g} s of the same recording, a1..a8, one core "
                         f"(the reference is single-threaded); {sum(ts):.2f} s in all"}

    stream_rec, fullrate_rec, sweep_rec = None, None, None
    was_graphed = graphed is not None
    if not args.no_extras and args.filter_mode == "parity":
        del A, Bn, graphed
        torch.cuda.empty_cache()
        fullrate_rec = fullrate_record(args, rank, world, dev, params, pcm, sr, max_over_ranks, barrier)
        torch.cuda.empty_cache()
        sw = sweep_record(args, rank, world, params)
        barrier()
        sw_ms = max_over_ranks(sw["ms_this_rank"])
        if rank == 0:
            hours = sw["settings"] * 0.5
            sweep_rec = {"workload": "C5: 256 band-pass / noise-floor settings over one synthetic 30-min 48 kHz recording, "
                                     f"settings sharded over {world} rank(s)", "ms_per_sweep": sw_ms,
                         "value": hours / (sw_ms / 1e3), "unit": UNIT, "settings": sw["settings"],
                         "rejected_like_the_reference_rank0": sw["rejected_like_the_reference"],
                         "timing": "wall clock around the whole sweep (one host synchronisation at its end), max over ranks"}
        torch.cuda.empty_cache()
    if not args.no_stream and not args.no_extras:
        torch.cuda.empty_cache()
        stream_rec = stream_record(args, rank, world, dev, params)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args), "filter_mode": args.filter_mode,
                           "raw_samples": len(pcm), "envelope_samples": M, "beats": len(beat_idx),
                           "l2": "inputs larger than L2 (PCM %.1f MB per step)" % (len(pcm) * 2 / 1e6),
                           "parallelism": f"{world} x independent recordings, no collective",
                           "host_cores": len(os.sched_getaffinity(0)),
                           **({"l2_fetch_granularity": int(l2_fetch)} if l2_fetch else {})},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h),
                        "api": "the drop-in functions: frontend.preprocess_pcm, _calculate_dynamic_noise_floor, "
                               "_initialize_state x2, calculate_bpm_series x2, find_recovery_phase, find_major_hr_inclines/"
                               "declines, calculate_hrr, find_peak_recovery/exertion_rate, calculate_windowed_hrv -- one "
                               "recording at a time, host arrays in, numpy / pandas objects out, sessions emptied every step",
                        "ingest": "bpm_host_gather_frames (host cores pack x[::ds] into pinned staging) + one cudaMemcpyAsync; "
                                  "results come back into pinned host arrays handed to the caller"},
                "e2e_pipelined": {"value": world * audio_hours / (pipe_ms / 1e3), "unit": UNIT, "ms_per_step": pipe_ms,
                        "ingest_mode": args.e2e_ingest if zero_copy else "full copy",
                        "ingest_candidates_ms": ingest_ms if zero_copy else None,
                        "h2d_bytes_per_step": int(pipe_h2d), "d2h_bytes_per_step": int(pipe_d2h),
                        "api": "runtime.StageAPipeline (throughput API: several recordings in flight)",
                        "ingest": (PIPE_INGEST[args.e2e_ingest] + f"{args.e2e_depth}-deep pipeline ingest | compute | "
                                   "read-back over three streams") if zero_copy else
                                  "cudaMemcpyAsync of the whole recording from pinned host memory, serial steps"},
                "gpu_launches": launches_per_step * args.steps if was_graphed else launches,
                "launch_mode": "cuda-graph replay" if was_graphed else "eager", "roofline": roofline,
                "cpu_baseline": cpu, "parity": parity, "fullrate": fullrate_rec, "sweep": sweep_rec,
                "stream": stream_rec, "kernels": kernels}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def fullrate_record(args, rank: int, world: int, dev, params, pcm, sr, max_over_ranks, barrier) -> dict:
    """The north star's headline filter ("full-rate Butterworth SOS band-pass", then decimation) on the same
    recording, device-resident, a1..a4: every one of the N raw samples is read and contracted
    (k_contract_i16), where the reference's own order (decimate first) reads one frame in `ds`."""
    import torch
    from bpm_analysis_b200.runtime import GraphedStep, StageARunner, profile_kernels
    fp = dict(params, filter_mode="fullrate")
    A = StageARunner([len(pcm)], sr, fp, want_filtered=True)
    A.upload([pcm])
    A.launch()
    torch.cuda.synchronize()
    g = None if args.no_graph else GraphedStep(A)
    step = A.launch if g is None else g.launch
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    rec = {"filter_mode": "fullrate", "scope": "a1..a4, device-resident", "ms_per_step": ms,
           "value": world * (len(pcm) / sr / 3600.0) / (ms / 1e3), "unit": UNIT, "raw_samples": len(pcm),
           "envelope_samples": A.total_m}
    prof = profile_kernels(lambda: [A.launch() for _ in range(5)])
    torch.cuda.synchronize()
    peak, peak_src = measured_peak_gbs()
    shp = {"N": len(pcm), "M": A.total_m, "T": int(A.out["trough_count"][0]), "P": int(A.out["peak_count"][0]), "B": 0}
    name = "k_contract_i16"
    if name in prof:
        cnt, tms = prof[name]
        us = tms * 1e3 / cnt
        ab = algorithmic_bytes(name, shp, "fullrate")
        rec["roofline"] = {"kernel": name, "bound": "hbm", "achieved": round(ab / (us * 1e-6) / 1e9, 1), "peak": peak,
                           "unit": "GB/s", "frac": round(ab / (us * 1e-6) / 1e9 / peak, 4), "avg_launch_us": round(us, 2),
                           "alg_bytes": ab, "peak_source": peak_src,
                           "share_of_step": round(tms / (sum(v[1] for v in prof.values()) or 1.0), 4),
                           # the kernel is bound by the FP64 (tensor) pipe, not by HBM: 8 multiply-adds per 2-byte sample
                           "fp64": {"flops": 16.0 * len(pcm), "achieved_tflops": round(16.0 * len(pcm) / (us * 1e-6) / 1e12, 2),
                                    "peak_tflops": 37.2, "frac": round(16.0 * len(pcm) / (us * 1e-6) / 37.2e12, 4),
                                    "peak_source": "148 SMs x 64 FP64 lanes x 2 x 1.965 GHz (ncu: the DMMA sub-pipe's own "
                                                   "peak is one m8n8k4 per 4.1 cycles per SM)"},
                           "note": "mma.m8n8k4.f64 on the FP64 tensor pipe (k_contract_i16_mma); the pipe, not HBM, is the "
                                   "ceiling: 2.76 GFLOP per launch"}
    if rank == 0 and not args.no_cpu_baseline:
        # parity on a bounded sample: scipy's sosfiltfilt at the full rate over the first 5 minutes, compared
        # on the first 4 (the prefix's own end transient, rho^k < 1e-22 after 4 k samples, is far away)
        from scipy.signal import butter, sosfiltfilt
        n5, keep_sec = int(300 * sr), 240
        nyq = 0.5 * sr
        sos = butter(2, [float(fp.get("lowcut_hz", 20.0)) / nyq, float(fp.get("highcut_hz", 150.0)) / nyq], btype="band",
                     output="sos")
        y = sosfiltfilt(sos, pcm[:n5].astype(np.float64), padlen=15)          # oracle/ref_port.py, 'fullrate'
        ds = A.plan.stride * A.plan.block
        k = keep_sec * sr // ds
        got = A.out["filtered"][:k].cpu().numpy()
        want = y[::ds][:k]
        rec["parity_filtered_first_4min"] = {"rel_err": float(np.max(np.abs(got - want)) / np.max(np.abs(want))),
                                             "against": "scipy.signal.sosfiltfilt at the full rate, then [::ds]",
                                             "tolerance": 1e-6}
    del A, g
    return rec


def sweep_record(args, rank: int, world: int, params) -> dict:
    """BASELINE configs[4] (C5) in front of the driver: 256 band-pass / noise-floor settings over one 30-min
    recording, sharded over the ranks (no collective); parity of settings against the oracle is the job
    of `--workload sweep` and tests/test_configs_gpu.py."""
    import torch
    from bpm_analysis_b200 import sweep, synth
    settings = synth.c5_settings()
    pcm, sr, _ = synth.config_c5(seed=5, duration_sec=1800.0)
    res = sweep.run_sweep(pcm, sr, params, settings, rank=rank, world=world)
    torch.cuda.synchronize()
    n = 2
    t0 = time.perf_counter()
    for _ in range(n):
        res = sweep.run_sweep(pcm, sr, params, settings, rank=rank, world=world)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / n
    return {"ms_this_rank": ms, "settings_this_rank": len(res), "settings": len(settings),
            "rejected_like_the_reference": sum(1 for r in res if "error" in r)}


def stream_record(args, rank: int, world: int, dev, params) -> dict:
    """SURVEY 8e row 2 in front of the driver: ONE long recording (C4, 24 h at 4 kHz) time-chunked
    over the ranks of this run with every stage evaluated per chunk (stream.ShardedFrontEnd); NCCL
    carries digit histograms, a table of counters and the chunks' trough / peak / strength lists.
    Rank 0 also evaluates the whole recording unchunked on its one GPU: that is the parity check
    (lists exact) and the one-GPU time the chunked step is compared with, measured in this run."""
    import torch
    import torch.distributed as dist
    from bpm_analysis_b200 import _native, stream, synth
    from bpm_analysis_b200.runtime import StageARunner
    lib = _native.load_library()
    sr, dur = 4000, float(args.stream_hours) * 3600.0
    n = int(dur * sr)
    pcm = None
    if rank == 0:
        pcm = synth.config_c4(seed=4, duration_sec=dur)[0]
        whole = torch.from_numpy(pcm).to(dev)
    else:
        whole = torch.empty(n, dtype=torch.int16, device=dev)
    if world > 1:
        dist.broadcast(whole.view(torch.uint8), 0)    # benchmark set-up only: every rank then keeps its own slice
    comm, eng = stream.DistComm(), stream.DeviceEngine()
    fe = stream.ShardedFrontEnd(n, sr, params, comm, eng)
    f0, f1 = fe.frames()
    pcm_dev = whole[f0:f1].clone()
    pcm_pin = torch.empty(f1 - f0, dtype=torch.int16).pin_memory()
    pcm_pin.copy_(pcm_dev)
    del whole
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    steps = max(3, min(args.steps, 10))
    use_graph = not args.no_graph
    l0 = lib.bpm_launch_count()
    out = fe.run(pcm_dev, want_filtered=False)                   # eager: counts the kernels of a step
    torch.cuda.synchronize()
    launches_per_step = int(lib.bpm_launch_count() - l0)

    def step():
        return fe.run(pcm_dev, want_filtered=False, graph=use_graph)

    for _ in range(3):
        out = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    barrier()
    launches = launches_per_step * steps
    graphed = use_graph and fe._graph is not None and not fe._graph.get("failed")
    ms_step = max_over_ranks(e0.elapsed_time(e1) / steps)
    # end to end: this rank's frames from pinned host memory; the gathered lists and this rank's own
    # envelope / floor chunk back into pinned host memory
    keys = ("troughs", "peaks", "strength", "smoothed_dev", "envelope_core", "floor_core")
    host = {}

    def e2e_step():
        pcm_dev.copy_(pcm_pin, non_blocking=True)
        o = step()
        for k in keys:
            if k not in host or host[k].numel() < o[k].numel():
                host[k] = torch.empty(o[k].numel(), dtype=o[k].dtype).pin_memory()
            host[k][:o[k].numel()].copy_(o[k], non_blocking=True)
        torch.cuda.synchronize()
        return o

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
    d2h = int(sum(out[k].numel() * out[k].element_size() for k in keys))
    sharded = bool(out["sharded"])
    audio_hours = dur / 3600.0
    rec = {"workload": f"C4 as ONE stream: synthetic {args.stream_hours:g}-h 4 kHz Holter-style recording, time-chunked "
                       f"over {world} GPU(s), every stage per chunk + halo (stream.ShardedFrontEnd)",
           "n_gpus": world, "scaling": "strong", "steps": steps, "ms_per_step": ms_step,
           "value": audio_hours / (ms_step / 1e3), "unit": UNIT, "chunk_proofs_held": sharded,
           "raw_samples": n, "envelope_samples": fe.chunks.m, "halo_envelope_samples": fe.chunks.halo,
           "frames_this_rank": int(f1 - f0), "gpu_launches": launches,
           "launch_mode": ("first part (filter .. proofs, NCCL included) as one CUDA graph, list exchange eager"
                           if graphed else "eager"),
           "exchange": "all_reduce(2048-bin key histograms) x3 + all_gather(quantile bucket), all_gather(8 counters), "
                       "all_gather(kept troughs | peaks | strength) -- never the envelope",
           "e2e": {"value": audio_hours / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": int((f1 - f0) * 2), "d2h_bytes_per_step": d2h},
           "result": {"troughs": int(out["troughs"].numel()), "peaks": int(out["peaks"].numel())}}
    # rank 0: the same recording unchunked on one GPU -- parity and the time to beat
    if rank == 0:
        A = StageARunner([n], sr, params, want_filtered=False)
        A.upload([pcm])
        A.launch()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(3):
            A.launch()
        a1.record()
        torch.cuda.synchronize()
        one_ms = a0.elapsed_time(a1) / 3
        nt, npk = int(A.out["trough_count"][0]), int(A.out["peak_count"][0])
        c0, c1 = fe.chunks.core(0)

        def rel(a, b):
            return float((a - b).abs().max() / b.abs().max())

        rec["one_gpu_unchunked_ms"] = one_ms
        rec["speedup_vs_one_gpu_unchunked"] = one_ms / ms_step
        rec["parity_vs_unchunked"] = {
            "troughs": "exact" if torch.equal(out["troughs"], A.out["troughs"][:nt]) else "DIFFERENT",
            "peaks": "exact" if torch.equal(out["peaks"], A.out["peaks"][:npk]) else "DIFFERENT",
            "strength_rel": rel(out["strength"], A.out["strength"][:npk]) if out["strength"].numel() == npk else None,
            "smoothed_dev_rel": rel(out["smoothed_dev"], A.out["smoothed_dev"][:max(npk - 1, 0)])
            if out["smoothed_dev"].numel() == max(npk - 1, 0) else None,
            "envelope_chunk_rel": rel(out["envelope_core"], A.out["envelope"][c0:c1]),
            "floor_chunk_rel": rel(out["floor_core"], A.out["floor"][c0:c1])}
        del A
    return rec


def run_stream(args):
    """BASELINE configs[1] as it is worded: ONE 60-min stream, halo-chunked over the GPUs
    (bpm_analysis_b200/stream.py).  Strong scaling: total work is fixed, every rank uploads
    only the PCM frames of its chunk + halo; NCCL all-gathers the envelope chunks, the kept
    trough lists and the floor chunks.  a1..a4 only (the beat-list reductions are replicated
    work of a few hundred microseconds)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from bpm_analysis_b200 import _native, stream
    lib = _native.load_library()
    params = bench_params(args)
    pcm, sr, _ = make_recording(args, 0)                       # every rank: the SAME recording
    comm = stream.DistComm()
    eng = stream.DeviceEngine()
    fe = stream.ChunkedFrontEnd(len(pcm), sr, params, comm, eng)
    f0, f1 = fe.frames()
    pcm_pin = torch.from_numpy(pcm[f0:f1].copy()).pin_memory()
    pcm_dev = pcm_pin.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = None
    for _ in range(args.warmup):
        out = fe.run(pcm_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = lib.bpm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = fe.run(pcm_dev)
    e1.record()
    barrier()
    launches = int(lib.bpm_launch_count() - l0)
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    audio_hours = args.duration_sec / 3600.0
    value = audio_hours / (ms_step / 1e3)
    # end to end: this rank's frames from pinned host memory, the full result back on every rank
    host = {k: torch.empty_like(v, device="cpu").pin_memory() for k, v in out.items()}
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pcm_dev.copy_(pcm_pin, non_blocking=True)
        out = fe.run(pcm_dev)
        for k, v in out.items():
            if host[k].numel() != v.numel():
                host[k] = torch.empty_like(v, device="cpu").pin_memory()
            host[k].copy_(v, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C2 as ONE stream: synthetic 60-min 48 kHz mono int16 recording, "
                                       f"halo-chunked over {world} GPU(s)",
                           "filter_mode": args.filter_mode, "raw_samples": len(pcm), "envelope_samples": fe.chunks.m,
                           "halo_envelope_samples": fe.chunks.halo, "frames_this_rank": int(f1 - f0),
                           "l2": "every step re-reads its PCM slice (%.1f MB)" % ((f1 - f0) * 2 / 1e6),
                           "parallelism": f"time chunks over {world} ranks; all_gather(envelope, kept troughs, floor)"},
                "clocks": clocks,
                "e2e": {"value": audio_hours / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int((f1 - f0) * 2),
                        "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in out.values()))},
                "gpu_launches": launches, "launch_mode": "eager (host reads list lengths between stages)",
                "roofline": None, "cpu_baseline": None,
                "result": {"troughs": int(out["troughs"].numel()), "peaks": int(out["peaks"].numel())}}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries write to fd 1 (NCCL prints its version there) goes to stderr; the ONE
    JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def run_extra(args):
    """Extra data points at sizes where the envelope-rate kernels stream more than L2: C3 (a rank's
    share of the 1024-recording batch) and C4 (24-h Holter stream).  a1..a4 only, device-resident
    timing + per-kernel event table; the e2e number is a serial pinned upload of the
    whole batch + compute + read-back."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from bpm_analysis_b200 import _native, synth
    from bpm_analysis_b200.runtime import GraphedStep, StageARunner, profile_kernels
    lib = _native.load_library()
    params = bench_params(args)
    if args.workload == "batch":
        base = [synth.config_c3_item(16 * rank + i)[0] for i in range(8)]
        pcms = [base[i % 8] for i in range(128)]
        sr, name = 44100, "C3: 128 x synthetic 10-min 44.1 kHz recordings per GPU (8 distinct, tiled), one bpm_stage_a call"
    else:
        pcm, sr, _ = synth.config_c4(seed=4 + rank, duration_sec=86400.0)
        pcms, name = [pcm], "C4: synthetic 24-h 4 kHz Holter-style recording (bursts, dropouts), one per GPU"
    audio_hours = sum(len(p) for p in pcms) / sr / 3600.0
    A = StageARunner([len(p) for p in pcms], sr, params)
    A.upload(pcms)
    torch.cuda.synchronize()
    graphed = None if args.no_graph else GraphedStep(A)
    step = A.launch if graphed is None else graphed.launch

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    l0 = lib.bpm_launch_count()
    A.launch()
    torch.cuda.synchronize()
    launches_per_step = int(lib.bpm_launch_count() - l0)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    # end to end, serial steps: the batch leaves ONE pinned buffer -- as strided copy-engine copies of
    # the kept frames when the decimation is coarse enough for that to beat copying everything
    # (bpm_copy_frames moves ~0.7 G frames/s, a plain copy ~55 GB/s: worth it from ds >= 40 for int16)
    # -- then compute, then envelope + floor + counts + the (distance-bounded) trough / peak lists
    # device -> pinned host
    pin_in = torch.empty(A.total_in, dtype=torch.int16).pin_memory()
    off = 0
    for p_ in pcms:
        pin_in.numpy()[off:off + len(p_)] = p_
        off += len(p_)
    sparse = args.filter_mode == "parity" and A.plan.stride >= 40
    E = StageARunner([len(p) for p in pcms], sr, params, pregathered="ce") if sparse else A
    dist_samples = max(int(E.cfg.distance), 1)
    caps = [int(it["m"]) // dist_samples + 2 for it in E.items]
    host = {k: torch.empty_like(E.out[k], device="cpu").pin_memory()
            for k in ("envelope", "floor", "trough_count", "peak_count")}
    host_lists = {k: [torch.empty(c, dtype=torch.int64).pin_memory() for c in caps] for k in ("troughs", "peaks")}
    d2h = sum(h.numel() * h.element_size() for h in host.values()) + 2 * 8 * sum(caps)
    h2d = E.total_m * 2 if sparse else A.total_in * 2

    def e2e_step():
        if sparse:
            E.gather(pin_in)
            E.launch()
        else:
            A.upload_pinned(pin_in)
            step()
        for k, h in host.items():
            h.copy_(E.out[k], non_blocking=True)
        for k, hs in host_lists.items():
            for it, h in zip(E.items, hs):
                h.copy_(E.out[k][int(it["m_off"]):int(it["m_off"]) + h.numel()], non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 5))
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / n_e2e)
    for it, ht, hp, ct, cp in zip(E.items, host_lists["troughs"], host_lists["peaks"], host["trough_count"].tolist(),
                                  host["peak_count"].tolist()):
        assert ct <= ht.numel() and cp <= hp.numel(), "find_peaks distance bounds the list lengths"
    assert torch.equal(host["peak_count"], A.out["peak_count"].cpu()) and torch.equal(host["floor"], A.out["floor"].cpu())
    clocks = sampler.stop()
    M = A.total_m
    nt, npk = int(host["trough_count"].sum()), int(host["peak_count"].sum())
    prof = profile_kernels(lambda: [A.launch() for _ in range(args.steps)])
    torch.cuda.synchronize()
    shp = {"N": sum(len(p) for p in pcms), "M": M, "T": nt, "P": npk, "B": 0}
    peak, peak_src = measured_peak_gbs()
    total_ms = sum(v[1] for v in prof.values()) or 1.0
    kernels = {}
    for kname, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        avg_us = ms * 1e3 / cnt
        ab = algorithmic_bytes(kname, shp, args.filter_mode)
        kernels[kname] = {"launches_per_step": cnt / args.steps, "avg_us": round(avg_us, 3), "share": round(ms / total_ms, 4),
                          "alg_bytes": ab, "gbs": round(ab / (avg_us * 1e-6) / 1e9, 2) if avg_us > 0 else None,
                          "frac_of_hbm_peak": round(ab / (avg_us * 1e-6) / 1e9 / peak, 4) if avg_us > 0 else None}
    top = next(iter(kernels))
    k = kernels[top]
    if rank == 0:
        emit({"metric": METRIC, "value": world * audio_hours / (ms_step / 1e3), "unit": UNIT, "n_gpus": world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": name, "filter_mode": args.filter_mode, "raw_samples": shp["N"],
                         "envelope_samples": M, "recordings": len(pcms), "troughs": nt, "raw_peaks": npk,
                         "l2": "inputs larger than L2", "parallelism": f"{world} x independent shares, no collective",
                         "scope": "a1..a4 (no beat-list reductions)"},
              "clocks": clocks,
              "e2e": {"value": world * audio_hours / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                      "ingest": ("serial steps: one pinned buffer -> device (" +
                                 ("bpm_copy_frames: strided copy-engine copies of the kept frames x[::ds]" if sparse else
                                  "cudaMemcpyAsync of the whole batch") +
                                 "), compute, envelope + floor + counts + distance-bounded lists -> pinned host")},
              "gpu_launches": launches_per_step * args.steps,
              "launch_mode": "eager" if graphed is None else "cuda-graph replay",
              "roofline": {"kernel": top, "bound": "hbm", "achieved": k["gbs"], "peak": peak, "unit": "GB/s",
                           "frac": k["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
                           "share_of_step": k["share"], "avg_launch_us": k["avg_us"], "note": ROOFLINE_NOTES.get(top, "")},
              "cpu_baseline": None, "kernels": kernels})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_sweep_bench(args):
    """BASELINE configs[4] ("C5"): 256 settings (16 band-passes x 16 noise-floor settings) over ONE
    30-min 48 kHz recording resident on the GPU; settings are independent units, sharded over the
    ranks with no collective.  A step = the whole sweep (a1 once per band-pass, a2..a4 per setting);
    value = settings x audio-hours / time.  Parity: 10 settings against the CPU oracle, in the run."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from bpm_analysis_b200 import _native, sweep, synth
    lib = _native.load_library()
    params = bench_params(args)
    settings = synth.c5_settings()
    pcm, sr, _ = synth.config_c5(seed=5, duration_sec=1800.0)
    audio_hours = len(settings) * (len(pcm) / sr / 3600.0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(1, args.warmup)):
        res = sweep.run_sweep(pcm, sr, params, settings, rank=rank, world=world)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = lib.bpm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        res = sweep.run_sweep(pcm, sr, params, settings, rank=rank, world=world)
    e1.record()
    barrier()
    wall_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
    ms_step = max_over_ranks(e0.elapsed_time(e1) / steps)
    launches = int(lib.bpm_launch_count() - l0)
    clocks = sampler.stop()
    parity, cpu = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref_port
        pick = [0, 17, 63, 64, 100, 150, 201, 202, 240, 255]
        got = sweep.run_sweep(pcm, sr, params, [settings[i] for i in pick], keep_arrays=True)
        ok, ts, errs = True, [], 0
        for rec, i in zip(got, pick):
            p = dict(params, **settings[i])
            if "error" in rec:
                errs += 1
                continue
            t1 = time.perf_counter()
            o = ref_port.front_end(pcm, sr, p)
            ts.append(time.perf_counter() - t1)
            e_env = float(np.max(np.abs(rec["envelope"].cpu().numpy() - o["envelope"])) / np.max(np.abs(o["envelope"])))
            e_fl = float(np.max(np.abs(rec["floor"].cpu().numpy() - o["floor"])) / np.max(np.abs(o["floor"])))
            ok = ok and e_env <= 1e-9 and e_fl <= 1e-9 and np.array_equal(rec["troughs"].cpu().numpy(), o["troughs"]) \
                and np.array_equal(rec["peaks"].cpu().numpy(), o["peaks"])
        parity = {"settings_checked": len(pick) - errs, "settings_rejected_like_the_reference": errs, "ok": bool(ok),
                  "tolerance": 1e-9}
        n_ok = sum(1 for r in res if "error" not in r)
        cpu = {"value": (len(pcm) / sr / 3600.0) / (sum(ts) / len(ts)), "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{len(ts)} of the {n_ok} accepted settings, a1..a4 each, one core; {sum(ts):.1f} s in all "
                         "(the reference has no sweep: it would run the whole path once per setting)"}
    if rank == 0:
        emit({"metric": METRIC, "value": audio_hours / (ms_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps,
              "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": "C5: 256 band-pass / noise-floor settings over one synthetic 30-min 48 kHz recording",
                         "filter_mode": args.filter_mode, "settings": len(settings),
                         "settings_rejected": sum(1 for r in res if "error" in r) if world == 1 else None,
                         "l2": "inputs larger than L2 (PCM 172.8 MB, read once per band-pass)",
                         "parallelism": f"settings sharded over {world} rank(s), no collective; "
                                        f"{sweep.N_STREAMS} streams per rank, one host synchronisation per sweep"},
              "clocks": clocks,
              "e2e": {"value": audio_hours / (wall_ms / 1e3), "unit": UNIT, "ms_per_step": wall_ms,
                      "h2d_bytes_per_step": int(len(pcm) * 2), "d2h_bytes_per_step": int(len(res) * 32),
                      "api": "sweep.run_sweep: host PCM array in (pageable), per-setting counts out; arrays stay on "
                             "the device"},
              "gpu_launches": launches, "launch_mode": "eager, 4 streams", "roofline": None, "cpu_baseline": cpu,
              "parity": parity})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "stream":
        run_stream(args)
    elif args.workload in ("batch", "holter"):
        run_extra(args)
    elif args.workload == "sweep":
        run_sweep_bench(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
