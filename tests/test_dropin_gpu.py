"""The session layer behind the drop-in functions (bpm_analysis_b200/dropin.py): results handed
from one reference call to the next are served from the device-side session, arrays that did not
come out of a session take the generic path, and stale or edited arrays are never trusted."""
import numpy as np
import pytest

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def svc():
    import torch
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    from bpm_analysis_b200 import _native, dropin
    _native.load_library()
    return dropin.dropin()


class _Clf:
    def __init__(self, env, rate, params):
        self.audio_envelope, self.sample_rate, self.params = env, rate, params


def test_chain_is_served_from_one_stage_a_call(svc, ref_params, synth_inputs):
    from bpm_analysis_b200 import frontend
    from oracle import ref_port
    pcm, sr = synth_inputs["c2_240s"]
    svc.forget()
    s0 = dict(svc.stats)
    env, rate, filt, dbg = frontend.preprocess_pcm(pcm, sr, ref_params, want_debug=True)
    floor, troughs = frontend._calculate_dynamic_noise_floor(env, rate, ref_params)
    clf = _Clf(env, rate, ref_params)
    st1 = frontend._initialize_state(clf, None, floor, troughs)
    peaks_only = frontend._find_raw_peaks(clf, floor.values)
    st2 = frontend._initialize_state(clf, 75.0, floor, troughs)
    s1 = dict(svc.stats)
    assert s1["stage_a_calls"] - s0["stage_a_calls"] == 1 and s1["session_misses"] == s0["session_misses"]
    o = ref_port.front_end(pcm, sr, ref_params)
    assert rel_err(env, o["envelope"]) < TOL and rel_err(filt, o["filtered"]) < TOL
    assert rel_err(floor.values, o["floor"]) < TOL
    assert np.array_equal(troughs, o["troughs"]) and troughs.dtype == np.int64
    for p in (st1["all_peaks"], st2["all_peaks"], peaks_only):
        assert np.array_equal(p, o["peaks"])
    assert st1["all_peaks"] is not st2["all_peaks"]                  # callers own what they get
    assert rel_err(st2["smoothed_dev_series"].values, o["smoothed_dev_series"].values) < TOL
    assert st2["long_term_bpm"] == 75.0 and st1["long_term_bpm"] == 80.0
    assert dbg.dtype == np.int16 and len(dbg) == len(env)


def test_foreign_arrays_take_the_generic_path_and_stay_resident(svc, ref_params):
    """The reference's own envelope (a plain numpy array): uploaded once for the three calls."""
    from bpm_analysis_b200 import frontend
    g = load_golden("vulpine")
    env, rate = g["envelope"].copy(), int(g["rate"])
    svc.forget()
    s0 = dict(svc.stats)
    floor, troughs = frontend._calculate_dynamic_noise_floor(env, rate, ref_params)
    clf = _Clf(env, rate, ref_params)
    peaks = frontend._find_raw_peaks(clf, floor.values)
    st = frontend._initialize_state(clf, None, floor, troughs)
    s1 = dict(svc.stats)
    assert s1["session_misses"] - s0["session_misses"] == 1 and s1["stage_a_calls"] == s0["stage_a_calls"]
    assert np.array_equal(troughs, g["troughs"]) and np.array_equal(peaks, g["raw_peaks"])
    assert np.array_equal(st["all_peaks"], g["raw_peaks"]) and rel_err(floor.values, g["floor"]) < TOL


def test_edited_or_recycled_arrays_are_not_trusted(svc, ref_params, synth_inputs):
    from bpm_analysis_b200 import frontend
    from oracle import ref_port
    pcm, sr = synth_inputs["c1_30s"]
    svc.forget()
    env, rate, _, _ = frontend.preprocess_pcm(pcm, sr, ref_params, want_filtered=False)
    floor, troughs = frontend._calculate_dynamic_noise_floor(env, rate, ref_params)
    # the caller scales its envelope in place: the session must notice and recompute from the new content
    env *= 2.0
    floor2, troughs2 = frontend._calculate_dynamic_noise_floor(env, rate, ref_params)
    o_floor, o_troughs = ref_port.calculate_dynamic_noise_floor(env, rate, ref_params)
    assert np.array_equal(troughs2, o_troughs) and rel_err(floor2.values, o_floor.values) < TOL
    # a different floor for the same envelope (e.g. a caller's own threshold array)
    clf = _Clf(env, rate, ref_params)
    other = np.full(len(env), float(np.quantile(env, 0.5)))
    pk = frontend._find_raw_peaks(clf, other)
    assert np.array_equal(pk, ref_port.find_raw_peaks(env, rate, ref_params, other))
    # different parameters for the same arrays
    p2 = dict(ref_params, noise_window_sec=4, noise_floor_quantile=0.3, peak_prominence_quantile=0.25)
    f3, t3 = frontend._calculate_dynamic_noise_floor(env, rate, p2)
    o3, ot3 = ref_port.calculate_dynamic_noise_floor(env, rate, p2)
    assert np.array_equal(t3, ot3) and rel_err(f3.values, o3.values) < TOL
    pk3 = frontend._find_raw_peaks(_Clf(env, rate, p2), f3.values)
    assert np.array_equal(pk3, ref_port.find_raw_peaks(env, rate, p2, o3.values))


def test_beat_list_reductions_share_one_round_trip(svc, ref_params, synth_inputs):
    from bpm_analysis_b200 import frontend, synth
    from oracle import ref_port
    pcm, sr, beats = synth.config_c2(seed=9, duration_sec=600.0)
    bi = synth.beats_to_envelope_indices(beats, 301)
    svc.forget()
    s0 = dict(svc.stats)
    sm, bt = frontend.calculate_bpm_series(bi, 301, ref_params)
    inc, dec = frontend.find_major_hr_inclines(sm), frontend.find_major_hr_declines(sm)
    rec, exe = frontend.find_peak_recovery_rate(sm), frontend.find_peak_exertion_rate(sm)
    hrv = frontend.calculate_windowed_hrv(bi, 301, ref_params)
    s1 = dict(svc.stats)
    assert s1["beat_misses"] - s0["beat_misses"] == 1 and s1["beat_hits"] - s0["beat_hits"] == 1
    o = ref_port.beat_reductions(bi, 301, ref_params)
    assert sm.index.equals(o["smoothed_bpm"].index) and rel_err(sm.values, o["smoothed_bpm"].values) < TOL
    assert [(x["start_time"], x["end_time"]) for x in inc] == [(x["start_time"], x["end_time"]) for x in o["major_inclines"]]
    assert [(x["start_time"], x["end_time"]) for x in dec] == [(x["start_time"], x["end_time"]) for x in o["major_declines"]]
    # the reductions of the Series are exact functions of ITS values: compare with the oracle on the same Series
    o_rec, o_exe = ref_port.find_peak_recovery_rate(sm), ref_port.find_peak_exertion_rate(sm)
    for got, want in ((rec, o_rec), (exe, o_exe)):
        assert got["slope_bpm_per_sec"] == want["slope_bpm_per_sec"] and got["duration_sec"] == want["duration_sec"]
        assert got["start_time"] == want["start_time"] and got["end_time"] == want["end_time"]
    assert rec["slope_bpm_per_sec"] == pytest.approx(o["peak_recovery_stats"]["slope_bpm_per_sec"], rel=1e-12)
    assert rel_err(hrv.values, o["windowed_hrv_df"].values) < TOL
    # a Series that did not come from calculate_bpm_series (shifted copy) takes the generic kernels
    other = sm * 1.0 + 0.25
    r2 = frontend.find_peak_recovery_rate(other)
    assert r2["slope_bpm_per_sec"] == ref_port.find_peak_recovery_rate(other)["slope_bpm_per_sec"]
    # non-default arguments as well
    r3 = frontend.find_peak_exertion_rate(sm, window_sec=35)
    assert r3["slope_bpm_per_sec"] == ref_port.find_peak_exertion_rate(sm, window_sec=35)["slope_bpm_per_sec"]
    i4 = frontend.find_major_hr_inclines(sm, min_duration_sec=30)
    assert [(x["start_time"], x["end_time"]) for x in i4] == \
        [(x["start_time"], x["end_time"]) for x in ref_port.find_major_hr_inclines(sm, min_duration_sec=30)]


def test_hrr_and_recovery_phase_mirrors(svc, ref_params):
    from bpm_analysis_b200 import frontend, synth
    from oracle import ref_port
    _, _, beats = synth.config_c2(seed=3, duration_sec=900.0)
    bi = synth.beats_to_envelope_indices(beats, 301)
    sm, bt = frontend.calculate_bpm_series(bi, 301, ref_params)
    got, want = frontend.calculate_hrr(sm), ref_port.calculate_hrr(sm)
    assert (got is None) == (want is None)
    if got is not None:
        assert got.keys() == want.keys() and all(got[k] == want[k] for k in got)
    assert frontend.find_recovery_phase(sm, bt, ref_params) == ref_port.find_recovery_phase(sm, bt, ref_params)
    assert frontend.find_recovery_phase(sm, bt[:1], ref_params) == (None, None)


def test_float32_mode_halves_the_read_back_and_keeps_the_lists(svc, ref_params, synth_inputs):
    """output_dtype="float32" (north star: "within ... a stated 1e-4 (float32 mode)"): signals come
    back as float32 within 1e-4 (measured ~6e-8) of the float64 reference values, every index list
    is identical to float64 mode (the device computes in float64 throughout), and the chain is
    still served from the one stage-A call."""
    from bpm_analysis_b200 import frontend
    from oracle import ref_port
    pcm, sr = synth_inputs["c2_240s"]
    p32 = dict(ref_params, output_dtype="float32")
    o = ref_port.front_end(pcm, sr, ref_params)
    svc.forget()
    s0 = dict(svc.stats)
    env, rate, filt, _ = frontend.preprocess_pcm(pcm, sr, p32)
    floor, troughs = frontend._calculate_dynamic_noise_floor(env, rate, p32)
    st = frontend._initialize_state(_Clf(env, rate, p32), None, floor, troughs)
    s1 = dict(svc.stats)
    assert s1["stage_a_calls"] - s0["stage_a_calls"] == 1 and s1["session_misses"] == s0["session_misses"]
    assert env.dtype == np.float32 and filt.dtype == np.float32 and floor.values.dtype == np.float32
    assert st["smoothed_dev_series"].values.dtype == np.float32
    assert rel_err(env.astype(np.float64), o["envelope"]) < 1e-4 and rel_err(env.astype(np.float64), o["envelope"]) < 1e-6
    assert rel_err(floor.values.astype(np.float64), o["floor"]) < 1e-4
    assert rel_err(st["smoothed_dev_series"].values.astype(np.float64), o["smoothed_dev_series"].values) < 1e-4
    assert np.array_equal(troughs, o["troughs"]) and np.array_equal(st["all_peaks"], o["peaks"])
    with pytest.raises(ValueError, match="output_dtype"):
        frontend.preprocess_pcm(pcm, sr, dict(ref_params, output_dtype="float16"))


def test_preprocess_audio_reads_the_wav_through_a_memory_map(svc, ref_params, synth_inputs, tmp_path):
    """File ingest: mono / stereo int16, float32 and uint8 WAVs are mapped, not
    copied; 24-bit PCM (which scipy cannot map) is mapped as bytes and only its kept frames are expanded
    (wav24.py, bpm_host_gather_s24).  Same envelope as the array-level call on the decoded samples."""
    import wave
    from scipy.io import wavfile
    from bpm_analysis_b200 import frontend
    from oracle import ref_port
    for name in ("c1_30s", "stereo_20s", "f32_20s", "u8_20s"):
        pcm, sr = synth_inputs[name]
        path = str(tmp_path / f"{name}.wav")
        wavfile.write(path, sr, pcm)
        _, mapped = frontend.read_wav(path)
        assert isinstance(mapped, np.memmap) and np.array_equal(np.asarray(mapped), pcm)
        env, rate = frontend.preprocess_audio(path, ref_params, str(tmp_path))
        o_env, o_rate, _ = ref_port.preprocess_pcm(pcm, sr, ref_params)
        assert rate == o_rate and rel_err(env, o_env) < TOL
    pcm, sr = synth_inputs["c1_30s"]
    path24 = str(tmp_path / "pcm24.wav")
    with wave.open(path24, "wb") as w:
        w.setnchannels(1), w.setsampwidth(3), w.setframerate(sr)
        w.writeframes((pcm.astype(np.int32) << 8).astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3].tobytes())
    from bpm_analysis_b200 import wav24
    sr24, x24 = frontend.read_wav(path24)
    assert sr24 == sr and not isinstance(x24, np.memmap) and isinstance(x24, wav24.S24Recording)
    env24, _ = frontend.preprocess_audio(path24, ref_params, str(tmp_path))
    o24, _, _ = ref_port.preprocess_pcm(wavfile.read(path24)[1], sr, ref_params)
    assert rel_err(env24, o24) < TOL
    # stereo, and a decimation small enough that the whole recording is expanded (no sparse gather)
    st, sr_st = synth_inputs["stereo_20s"]
    path24s = str(tmp_path / "pcm24s.wav")
    with wave.open(path24s, "wb") as w:
        w.setnchannels(2), w.setsampwidth(3), w.setframerate(sr_st)
        w.writeframes((st.astype(np.int32) << 8).astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3].tobytes())
    for p in (ref_params, dict(ref_params, downsample_factor=2)):
        env_s, rate_s = frontend.preprocess_audio(path24s, p, str(tmp_path))
        o_s, o_rate_s, _ = ref_port.preprocess_pcm(wavfile.read(path24s)[1], sr_st, p)
        assert rate_s == o_rate_s and rel_err(env_s, o_s) < (TOL if rate_s <= 12000 else 1e-6)
