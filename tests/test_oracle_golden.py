"""The oracle (oracle/ref_port.py) against what the UNMODIFIED reference produced.

Fixtures: tests/golden/*.npz, written by oracle/make_golden.py in the authoring
container from /root/reference (its shipped samples/ run and live calls into its
module).  These tests run on CPU and pin the oracle before any GPU test trusts it.
"""
import hashlib

import numpy as np
import pandas as pd
import pytest

from conftest import load_golden, rel_err
from oracle import ref_port

SYNTH = ["c1_30s", "c2_240s", "holter_180s", "short_1s5", "stereo_20s", "f32_20s", "u8_20s", "i32_20s"]


def test_vulpine_log_pins_reference_run():
    """The reference's shipped Debug_Log pins raw peaks exactly; 1345/1349 troughs exactly.

    The 4 others moved (3 by one sample, 1 by nine) because the shipped WAV is the int16
    quantisation of the float signal the logged run saw."""
    g = load_golden("vulpine")
    assert np.array_equal(g["raw_peaks"], g["log_peak_idx"])
    assert len(g["troughs"]) == len(g["log_trough_idx"]) == 1349
    assert np.max(np.abs(g["troughs"] - g["log_trough_idx"])) <= 9
    assert np.sum(g["troughs"] != g["log_trough_idx"]) <= 4
    # BPM CSV (733 rows, %.3f)
    assert len(g["csv_bpm"]) == len(g["bpm_values"]) == 733
    assert np.max(np.abs(np.round(g["bpm_values"], 3) - g["csv_bpm"])) < 1.5e-3
    assert np.max(np.abs(np.round(g["bpm_times"], 3) - g["csv_time"])) < 1.5e-3


def test_vulpine_oracle_chain(ref_params):
    g = load_golden("vulpine")
    rate = int(g["rate"])
    env = ref_port.envelope_of(g["filtered_i16"].astype(np.float64), rate)
    assert np.array_equal(env, g["envelope"])
    floor, troughs = ref_port.calculate_dynamic_noise_floor(env, rate, ref_params)
    assert np.array_equal(troughs, g["troughs"])
    assert np.array_equal(floor.values, g["floor"])
    peaks = ref_port.find_raw_peaks(env, rate, ref_params, floor.values)
    assert np.array_equal(peaks, g["raw_peaks"])
    assert np.array_equal(peaks, g["log_peak_idx"])
    pm = ref_port.peak_metrics(env, rate, ref_params, floor, peaks)
    assert np.array_equal(pm["smoothed_dev_series"].values, g["smoothed_dev_values"])
    assert np.array_equal(pm["smoothed_dev_series"].index.values, g["smoothed_dev_index"])


def test_vulpine_oracle_beat_reductions(ref_params):
    g = load_golden("vulpine")
    rate = int(g["rate"])
    r = ref_port.beat_reductions(g["beats"], rate, ref_params)
    s = r["smoothed_bpm"]
    assert np.array_equal(s.values, g["bpm_values"])
    assert np.array_equal(s.index.as_unit("us").asi8, g["bpm_index_us"])
    assert np.array_equal(r["bpm_times"], g["bpm_times"])
    assert np.array_equal(r["windowed_hrv_df"][["time", "rmssdc", "sdnn", "bpm"]].values, g["hrv"])
    # summary numbers the reference shipped (vulpine_Analysis_Summary.md:6-30)
    avg, lo, hi, rmssdc, sdnn, ex_s, ex_d, re_s, re_d = g["summary"]
    assert round(s.mean(), 1) == avg and round(s.min(), 1) == lo and round(s.max(), 1) == hi
    assert round(r["windowed_hrv_df"]["rmssdc"].mean(), 2) == rmssdc
    assert round(r["windowed_hrv_df"]["sdnn"].mean(), 2) == sdnn
    assert round(r["peak_exertion_stats"]["slope_bpm_per_sec"], 2) == ex_s
    assert round(r["peak_exertion_stats"]["duration_sec"], 1) == ex_d
    assert round(r["peak_recovery_stats"]["slope_bpm_per_sec"], 2) == re_s
    assert round(r["peak_recovery_stats"]["duration_sec"], 1) == re_d
    assert len(r["major_inclines"]) == int(g["inclines_n"]) == 7
    assert len(r["major_declines"]) == int(g["declines_n"]) == 4
    for key, lst, ch in (("inclines", r["major_inclines"], "bpm_increase"),
                         ("declines", r["major_declines"], "bpm_decrease")):
        assert np.array_equal([d["slope_bpm_per_sec"] for d in lst], g[key + "_slope_bpm_per_sec"])
        assert np.array_equal([d[ch] for d in lst], g[key + "_" + ch])
        assert np.array_equal([pd.Timestamp(d["start_time"]).as_unit("us").value for d in lst],
                              g[key + "_start_us"])


@pytest.mark.parametrize("name", SYNTH)
def test_synth_oracle_matches_reference(name, ref_params, synth_inputs):
    g = load_golden("synth_" + name)
    pcm, sr = synth_inputs[name]
    assert hashlib.sha256(np.ascontiguousarray(pcm).tobytes()).hexdigest() == str(g["pcm_sha256"]), \
        "synthetic generator drifted from the fixture"
    env, rate, filt = ref_port.preprocess_pcm(pcm, sr, ref_params)
    assert rate == int(g["rate"])
    assert np.array_equal(env, g["envelope"])
    assert np.array_equal(ref_port.debug_wav_samples(filt), g["debug_i16"])
    floor, troughs = ref_port.calculate_dynamic_noise_floor(env, rate, ref_params)
    assert np.array_equal(np.asarray(troughs, dtype=np.int64), g["troughs"])
    assert np.array_equal(floor.values, g["floor"])
    peaks = ref_port.find_raw_peaks(env, rate, ref_params, floor.values)
    assert np.array_equal(peaks, g["raw_peaks"])
    if len(peaks) >= 2:
        pm = ref_port.peak_metrics(env, rate, ref_params, floor, peaks)
        assert np.array_equal(pm["smoothed_dev_series"].values, g["smoothed_dev_values"])
    beats = g["beats"]
    if len(beats) >= 2:
        r = ref_port.beat_reductions(beats, rate, ref_params)
        assert np.array_equal(r["smoothed_bpm"].values, g["bpm_values"])
        assert np.array_equal(r["smoothed_bpm"].index.as_unit("us").asi8, g["bpm_index_us"])
        hrv = r["windowed_hrv_df"]
        got = hrv[["time", "rmssdc", "sdnn", "bpm"]].values if len(hrv) else np.zeros((0, 4))
        assert np.array_equal(np.asarray(got, dtype=np.float64), g["hrv"])
        for key, d in (("recovery", r["peak_recovery_stats"]), ("exertion", r["peak_exertion_stats"])):
            assert int(g[key + "_present"]) == (0 if d is None else 1)
            if d is not None:
                assert d["slope_bpm_per_sec"] == float(g[key + "_slope_bpm_per_sec"])
                assert pd.Timestamp(d["start_time"]).as_unit("us").value == int(g[key + "_start_us"])
        assert len(r["major_inclines"]) == int(g["inclines_n"])
        assert len(r["major_declines"]) == int(g["declines_n"])


@pytest.mark.parametrize("name", ["vulpine"] + ["synth_" + n for n in SYNTH])
def test_hrr_and_recovery_phase_match_reference(name, ref_params):
    """calculate_hrr / find_recovery_phase (bpm_analysis.py:1597-1620) as the unmodified reference
    evaluated them on its own final beat list -- including what the installed pandas does with the
    index arithmetic of :1604-1607 (vulpine: 75.5, not the 58.9 of the shipped summary)."""
    g = load_golden(name)
    if len(g["beats"]) < 2:
        pytest.skip("no beat series in this fixture")
    r = ref_port.beat_reductions(g["beats"], int(g["rate"]), ref_params)
    hrr = r["hrr_stats"]
    assert (hrr is not None) == bool(g["hrr_found"])
    if hrr is not None:
        assert np.array_equal([hrr["peak_bpm"], hrr["recovery_bpm"], hrr["hrr_value_bpm"], hrr["interval_sec"]],
                              g["hrr_values"])
        assert [pd.Timestamp(hrr["peak_time"]).as_unit("us").value,
                pd.Timestamp(hrr["recovery_check_time"]).as_unit("us").value] == list(g["hrr_times_us"])
    want = [None if np.isnan(v) else float(v) for v in g["recovery_phase"]]
    assert list(r["recovery_phase"]) == want
