"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.npz``.

Run in the authoring container (needs ``/root/reference``):

    python -m oracle.make_golden

Two kinds of fixture:

* ``vulpine.npz`` -- the reference's own shipped run (``samples/vulpine_*``),
  re-encoded: the 302 Hz int16 band-passed signal, every trough / raw-peak time
  in ``vulpine_Debug_Log.md`` as an index, the BPM-series CSV, the summary
  numbers -- plus what the unmodified reference module computes from that
  signal here (floor, troughs, raw peaks, beat list, HRV, slopes).
* ``synth_*.npz`` -- outputs of the unmodified reference on small seeded
  synthetic recordings (``bpm_analysis_b200.synth``).  Inputs are re-generated
  from the seed at test time (a SHA-256 of the PCM is stored to catch drift),
  so the fixtures stay small.
"""
from __future__ import annotations

import hashlib
import os
import re
import sys
import tempfile

import numpy as np
import pandas as pd
from scipy.io import wavfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from bpm_analysis_b200 import synth                       # noqa: E402
from oracle.load_reference import REFERENCE_ROOT, load_reference, reference_params  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")


def _series_pack(prefix, s: pd.Series, out):
    out[prefix + "_values"] = np.asarray(s.values, dtype=np.float64)
    if isinstance(s.index, pd.DatetimeIndex):
        out[prefix + "_index_us"] = s.index.as_unit("us").asi8.copy()
        out[prefix + "_index_unit"] = np.array(str(s.index.dtype))
    else:
        out[prefix + "_index"] = np.asarray(s.index.values)


def _period_pack(prefix, d, out, epoch_us):
    if d is None:
        out[prefix + "_present"] = np.array(0)
        return
    out[prefix + "_present"] = np.array(1)
    out[prefix + "_start_us"] = np.array(pd.Timestamp(d["start_time"]).as_unit("us").value)
    out[prefix + "_end_us"] = np.array(pd.Timestamp(d["end_time"]).as_unit("us").value)
    for k in ("start_bpm", "end_bpm", "slope_bpm_per_sec", "duration_sec"):
        out[prefix + "_" + k] = np.array(float(d[k]))


def _segments_pack(prefix, lst, out, change_key):
    out[prefix + "_n"] = np.array(len(lst))
    for k in ("start_bpm", "end_bpm", "duration_sec", change_key, "slope_bpm_per_sec"):
        out[prefix + "_" + k] = np.array([float(d[k]) for d in lst], dtype=np.float64)
    out[prefix + "_start_us"] = np.array([pd.Timestamp(d["start_time"]).as_unit("us").value for d in lst],
                                         dtype=np.int64)
    out[prefix + "_end_us"] = np.array([pd.Timestamp(d["end_time"]).as_unit("us").value for d in lst],
                                       dtype=np.int64)


def run_reference_chain(ref, env, rate, params, out):
    """The numeric stages of analyze_wav_file (:1731-1757) on a given envelope."""
    floor, troughs = ref._calculate_dynamic_noise_floor(env, rate, params)
    out["floor"] = np.asarray(floor.values, dtype=np.float64)
    out["troughs"] = np.asarray(troughs, dtype=np.int64)
    start_bpm, peak_t, rec_t = ref._run_preliminary_pass(env, rate, params, floor, troughs, None)
    clf = ref.PeakClassifier(env, rate, params, start_bpm, floor, troughs, peak_t, rec_t)
    st = clf.state
    out["raw_peaks"] = np.asarray(st["all_peaks"], dtype=np.int64)
    if len(st["all_peaks"]) >= 2:
        _series_pack("smoothed_dev", st["smoothed_dev_series"], out)
    s1, all_raw, adata = clf.classify_peaks()
    if len(s1) >= 2:
        final, adata = ref._refine_and_correct_peaks(s1, all_raw, adata, env, rate, params)
    else:
        final = np.asarray(s1)
    final = np.asarray(final, dtype=np.int64)
    out["beats"] = final
    if len(final) < 2:
        return
    series, times = ref.calculate_bpm_series(final, rate, params)
    _series_pack("bpm", series, out)
    out["bpm_times"] = np.asarray(times, dtype=np.float64)
    epoch_us = 0
    _period_pack("recovery", ref.find_peak_recovery_rate(series), out, epoch_us)
    _period_pack("exertion", ref.find_peak_exertion_rate(series), out, epoch_us)
    _segments_pack("inclines", ref.find_major_hr_inclines(series), out, "bpm_increase")
    _segments_pack("declines", ref.find_major_hr_declines(series), out, "bpm_decrease")
    # the two remaining helpers of the metrics chain (:1597-1620), as the reference evaluates them under
    # the pandas installed here (calculate_hrr divides a datetime64[us] index by 1e9)
    hrr = ref.calculate_hrr(series)
    out["hrr_found"] = np.array(hrr is not None)
    if hrr is not None:
        out["hrr_values"] = np.array([hrr["peak_bpm"], hrr["recovery_bpm"], hrr["hrr_value_bpm"], hrr["interval_sec"]],
                                     dtype=np.float64)
        out["hrr_times_us"] = np.array([pd.Timestamp(hrr["peak_time"]).as_unit("us").value,
                                        pd.Timestamp(hrr["recovery_check_time"]).as_unit("us").value], dtype=np.int64)
    phase = ref.find_recovery_phase(series, times, params)
    out["recovery_phase"] = np.array([np.nan if v is None else float(v) for v in phase], dtype=np.float64)
    hrv = ref.calculate_windowed_hrv(final, rate, params)
    out["hrv"] = np.asarray(hrv[["time", "rmssdc", "sdnn", "bpm"]].values, dtype=np.float64) \
        if len(hrv) else np.zeros((0, 4))


def golden_vulpine(ref, params):
    sdir = os.path.join(REFERENCE_ROOT, "samples")
    rate, y = wavfile.read(os.path.join(sdir, "vulpine_filtered_debug.wav"))
    out = {"rate": np.array(rate), "filtered_i16": y.astype(np.int16)}
    # what the reference's own log pins (times printed to 0.1 ms; index = round(t*rate))
    txt = open(os.path.join(sdir, "vulpine_Debug_Log.md"), encoding="utf-8").read()
    blocks = re.split(r"\n## Time: `([0-9.]+)s`\n", txt)
    log_troughs, log_peaks = [], []
    for k in range(1, len(blocks), 2):
        t, body = float(blocks[k]), blocks[k + 1]
        (log_troughs if body.lstrip().startswith("**Trough Detected**") else log_peaks).append(
            int(round(t * rate)))
    out["log_trough_idx"] = np.array(sorted(set(log_troughs)), dtype=np.int64)
    out["log_peak_idx"] = np.array(sorted(set(log_peaks)), dtype=np.int64)
    csv = pd.read_csv(os.path.join(sdir, "vulpine_bpm_plot.csv"))
    out["csv_time"] = csv.iloc[:, 0].values.astype(np.float64)
    out["csv_bpm"] = csv.iloc[:, 1].values.astype(np.float64)
    # summary numbers (vulpine_Analysis_Summary.md:6-30)
    out["summary"] = np.array([122.2, 78.6, 163.3, 117.97, 70.29, 3.35, 20.1, -3.11, 20.7])
    # envelope exactly as the reference's own second consumer of this WAV builds it
    # (heartbeat_labeler.py:63-67 == bpm_analysis.py:1052-1054)
    env = pd.Series(np.abs(y.astype(np.float64))).rolling(
        window=rate // 10, min_periods=1, center=True).mean().values
    out["envelope"] = env
    run_reference_chain(ref, env, int(rate), params, out)
    np.savez_compressed(os.path.join(GOLDEN, "vulpine.npz"), **out)
    print("vulpine:", {k: getattr(v, "shape", None) for k, v in out.items() if k in
                       ("raw_peaks", "troughs", "beats", "log_peak_idx", "log_trough_idx", "hrv")})


def golden_synth(ref, base_params, name, pcm, sr, extra_params=None):
    params = dict(base_params)
    params["save_filtered_wav"] = False
    if extra_params:
        params.update(extra_params)
    out = {"sample_rate": np.array(sr),
           "pcm_sha256": np.array(hashlib.sha256(np.ascontiguousarray(pcm).tobytes()).hexdigest())}
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "in.wav")
        wavfile.write(p, sr, pcm)
        # the filtered signal is only observable through the debug WAV; get it at f64 by
        # calling the same two scipy lines the reference calls (:1044-1045) -- checked
        # below against the int16 debug WAV the reference writes
        env, rate = ref.preprocess_audio(p, params, td)
        pdbg = dict(params)
        pdbg["save_filtered_wav"] = True
        ref.preprocess_audio(p, pdbg, td)
        _, dbg = wavfile.read(os.path.join(td, "in_filtered_debug.wav"))
    out["envelope"] = np.asarray(env, dtype=np.float64)
    out["rate"] = np.array(rate)
    out["debug_i16"] = dbg.astype(np.int16)
    run_reference_chain(ref, env, int(rate), params, out)
    np.savez_compressed(os.path.join(GOLDEN, f"synth_{name}.npz"), **out)
    print(name, "M=", len(env), "troughs", len(out["troughs"]), "peaks", len(out["raw_peaks"]),
          "beats", len(out["beats"]))


def synth_cases():
    """name -> (pcm, sr).  Kept in one place so tests regenerate identical inputs."""
    cases = {}
    cases["c1_30s"] = synth.config_c1(seed=11, duration_sec=30.0)[:2]
    cases["c2_240s"] = synth.config_c2(seed=12, duration_sec=240.0)[:2]
    pcm, sr, _ = synth.pcg_recording(180.0, 4000, lambda t: 72.0 + 8.0 * np.sin(t / 20.0), 13,
                                     bursts=[(40.0, 4.0)], dropouts=[(90.0, 3.0), (150.0, 1.5)])
    cases["holter_180s"] = (pcm, sr)
    cases["short_1s5"] = synth.pcg_recording(2.6, 44100, lambda t: 70.0, 14)[:2]
    pcm, sr, _ = synth.config_c1(seed=15, duration_sec=20.0)
    cases["stereo_20s"] = (np.stack([pcm, (pcm // 2).astype(np.int16)], axis=1), sr)
    pcm, sr, _ = synth.config_c1(seed=16, duration_sec=20.0)
    cases["f32_20s"] = ((pcm.astype(np.float32) / 32768.0), sr)
    pcm, sr, _ = synth.config_c1(seed=17, duration_sec=20.0)
    cases["u8_20s"] = (((pcm.astype(np.int32) >> 8) + 128).astype(np.uint8), sr)
    pcm, sr, _ = synth.config_c1(seed=18, duration_sec=20.0)
    cases["i32_20s"] = ((pcm.astype(np.int32) << 12), sr)
    return cases


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = load_reference()
    params = reference_params()
    golden_vulpine(ref, params)
    for name, (pcm, sr) in synth_cases().items():
        golden_synth(ref, params, name, pcm, sr)


if __name__ == "__main__":
    main()
