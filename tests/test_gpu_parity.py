"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the golden
fixtures the unmodified reference produced.  Tolerances: index outputs
bit-exact; float signals max|d| <= 1e-9 * max|ref|."""
import numpy as np
import pandas as pd
import pytest
from scipy.signal import find_peaks as scipy_find_peaks

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9
SYNTH = ["c1_30s", "c2_240s", "holter_180s", "short_1s5", "stereo_20s", "f32_20s", "u8_20s", "i32_20s"]


@pytest.fixture(scope="module")
def fe():
    import torch
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    from bpm_analysis_b200 import _native, frontend
    _native.load_library()          # fail loudly if the extension is missing
    return frontend


@pytest.fixture(scope="module")
def ops(fe):
    from bpm_analysis_b200 import runtime
    return runtime.ops()


# ----------------------------------------------------------------------------- a1
@pytest.mark.parametrize("name", SYNTH)
def test_preprocess_matches_reference_golden(name, fe, ref_params, synth_inputs):
    g = load_golden("synth_" + name)
    pcm, sr = synth_inputs[name]
    env, rate, filt, dbg = fe.preprocess_pcm(pcm, sr, ref_params, want_debug=True)
    assert rate == int(g["rate"])
    assert rel_err(env, g["envelope"]) < TOL
    # int16 debug WAV: truncating cast of y/max|y|*32767 -- identical except where the float
    # value sits within rounding of an integer
    d = dbg.astype(np.int32) - g["debug_i16"].astype(np.int32)
    assert np.max(np.abs(d)) <= 1 and np.mean(d != 0) < 1e-3


@pytest.mark.parametrize("name", ["c1_30s", "holter_180s", "stereo_20s"])
def test_preprocess_fullrate_matches_oracle(name, fe, ref_params, synth_inputs):
    from oracle import ref_port
    pcm, sr = synth_inputs[name]
    p = dict(ref_params, filter_mode="fullrate")
    env, rate, filt, _ = fe.preprocess_pcm(pcm, sr, p)
    o_env, o_rate, o_filt = ref_port.preprocess_pcm(pcm, sr, p)
    assert rate == o_rate
    assert rel_err(filt, o_filt) < TOL
    assert rel_err(env, o_env) < TOL


def test_preprocess_audio_file_roundtrip(fe, ref_params, synth_inputs, tmp_path):
    from scipy.io import wavfile
    pcm, sr = synth_inputs["c1_30s"]
    g = load_golden("synth_c1_30s")
    wav = tmp_path / "rec.wav"
    wavfile.write(str(wav), sr, pcm)
    out_dir = tmp_path / "out"
    out_dir.mkdir()
    p = dict(ref_params, save_filtered_wav=True)
    env, rate = fe.preprocess_audio(str(wav), p, str(out_dir))
    assert rate == 302 and rel_err(env, g["envelope"]) < TOL
    r1, d1 = wavfile.read(str(tmp_path / "rec_filtered_debug.wav"))
    r2, d2 = wavfile.read(str(out_dir / "rec_filtered_debug.wav"))
    assert r1 == r2 == 302 and d1.dtype == np.int16 and np.array_equal(d1, d2)
    assert np.max(np.abs(d1.astype(int) - g["debug_i16"].astype(int))) <= 1


def test_preprocess_errors_match_reference(fe, ref_params):
    with pytest.raises(ValueError, match="padlen"):
        fe.preprocess_pcm(np.zeros(15 * 146, dtype=np.int16), 44100, ref_params)
    with pytest.raises(KeyError):
        fe.preprocess_audio("/nonexistent.wav", {}, "/tmp")
    with pytest.raises(ValueError, match="filter"):
        fe.preprocess_pcm(np.zeros(5000, dtype=np.int16), 250, dict(ref_params, downsample_factor=1))


# ----------------------------------------------------------------------------- K3
@pytest.mark.parametrize("n,q", [(1, 0.1), (2, 0.5), (7, 0.2), (1000, 0.1), (90617, 0.1), (90617, 0.2),
                                 (300001, 0.37), (5000, 0.0), (5000, 1.0)])
def test_quantile_bit_exact(n, q, ops):
    rng = np.random.default_rng(n)
    x = np.abs(rng.standard_normal(n)) * 1000.0
    if n > 100:
        x[rng.integers(0, n, n // 3)] = x[0]          # heavy duplicates
        x[rng.integers(0, n, n // 50)] = 0.0
    assert ops.quantile(x, q) == np.quantile(x, q)
    y = rng.standard_normal(n)                        # negative values too
    assert ops.quantile(y, q) == np.quantile(y, q)


@pytest.mark.parametrize("q", [0.0, 0.1, 0.2, 0.5, 0.93, 1.0])
def test_quantile_large_buckets_slow_path(q, ops):
    """Buckets larger than the collect capacity (4096 keys sharing their leading 22 bits): exact
    duplicates straddling the target rank, a constant array, and distinct values packed into
    one binade slice -- all resolved by the finishing CTA's remaining digit passes."""
    rng = np.random.default_rng(17)
    a = np.abs(rng.standard_normal(60000)) * 10.0
    a[rng.integers(0, 60000, 25000)] = 3.25                 # ~20 k exact duplicates
    assert ops.quantile(a, q) == np.quantile(a, q)
    c = np.full(20000, 7.5)
    assert ops.quantile(c, q) == np.quantile(c, q)
    d = 1.0 + 1e-9 * rng.random(50000)                      # same sign, exponent and top mantissa bits
    assert ops.quantile(d, q) == np.quantile(d, q)
    e = np.concatenate([np.zeros(30000), -np.abs(rng.standard_normal(100)), np.abs(rng.standard_normal(5000))])
    assert ops.quantile(e, q) == np.quantile(e, q)


@pytest.mark.parametrize("q", [0.1, 0.2, 0.77])
def test_quantile_long_stream_regime(q, ops):
    """More than 4 M samples: three digit passes before the bucket is collected (the 24-h Holter
    stream has 28.8 M envelope samples); with and without a block of exact zeros (dropouts) that
    swallows the target rank."""
    rng = np.random.default_rng(23)
    x = np.abs(rng.standard_normal(5_000_000)) ** 2.5 * 300.0
    assert ops.quantile(x, q) == np.quantile(x, q)
    x[1_000_000:2_200_000] = 0.0                           # 24 % digital silence
    assert ops.quantile(x, q) == np.quantile(x, q)


# ----------------------------------------------------------------------------- K4
def _signals():
    rng = np.random.default_rng(5)
    out = {"noise": rng.standard_normal(20000)}
    s = np.round(rng.standard_normal(30000) * 3.0)     # many ties and plateaus
    out["quantised"] = s
    t = np.arange(50000) / 300.0
    out["ramp_ripple"] = t * 0.2 + 0.01 * np.sin(2 * np.pi * 40 * t) + 0.001 * rng.standard_normal(t.size)
    z = np.abs(rng.standard_normal(40000))
    z[10000:12000] = 0.0
    z[30000:30007] = 5.0
    z[-50:] = 0.0
    out["dropout"] = z
    out["tiny"] = np.array([0.0, 1.0, 0.0, 2.0, 2.0, 0.0, 3.0])
    return out


@pytest.mark.parametrize("name", ["noise", "quantised", "ramp_ripple", "dropout", "tiny"])
@pytest.mark.parametrize("distance", [1, 3, 15, 200])
def test_find_peaks_matches_scipy(name, distance, ops):
    x = _signals()[name]
    prom = float(np.quantile(np.abs(x), 0.3))
    for sign in (+1, -1):
        ref_np, _ = scipy_find_peaks(sign * x, distance=distance)
        ref_p, _ = scipy_find_peaks(sign * x, distance=distance, prominence=prom)
        got_np = ops.find_peaks(x, None, None, distance, sign)
        got_p = ops.find_peaks(x, None, prom, distance, sign)
        if name == "quantised" and distance > 1:
            # equal heights inside `distance`: scipy's survivor depends on an unstable sort;
            # only the count of survivors per tie-free region is pinned -> compare loosely
            assert abs(len(got_np) - len(ref_np)) <= 0.02 * len(ref_np) + 2
            continue
        assert np.array_equal(got_np, ref_np)
        assert np.array_equal(got_p, ref_p)


def test_find_peaks_long_plateaus(ops):
    """Flat runs of every interesting length around the 512-step walk bound (shorter, exactly at
    2*512+1, longer, far longer), as strict maxima, as shelves (one side higher) and touching the
    array ends; plus a constant signal of a million samples, which must simply return nothing."""
    rng = np.random.default_rng(9)
    parts, level = [], 0.0
    for L in (1, 2, 3, 5, 511, 512, 513, 1023, 1024, 1025, 1026, 1027, 2049, 5000, 40000):
        for kind in ("max", "shelf_up", "shelf_down", "min"):
            base = rng.random(40) * 0.1
            lo = float(base.max()) + 0.5
            run = np.full(L, lo + 1.0)
            left = base + (2.0 if kind == "shelf_down" else 0.0) + (3.0 if kind == "min" else 0.0)
            right = rng.random(40) * 0.1 + (2.0 if kind == "shelf_up" else 0.0) + (3.0 if kind == "min" else 0.0)
            parts += [left, run, right]
    x = np.concatenate([np.full(700, 9.0)] + parts + [np.full(1500, 9.0)])       # flat runs touching both ends
    for sign in (1, -1):
        ref, _ = scipy_find_peaks(sign * x)
        got = ops.find_peaks(x, sign=sign)
        assert np.array_equal(got, ref)
    ref, _ = scipy_find_peaks(x, distance=7, prominence=0.5)
    assert np.array_equal(ops.find_peaks(x, distance=7, prominence=0.5), ref)
    flat = np.full(1_000_000, 0.25)
    assert len(ops.find_peaks(flat)) == 0 and len(ops.find_peaks(flat, sign=-1)) == 0
    flat[400_000] = 0.2                                    # two 400 k / 600 k shelves around a notch: still no maximum
    assert len(ops.find_peaks(flat)) == 0
    assert np.array_equal(ops.find_peaks(flat, sign=-1), [400_000])
    bump = np.zeros(900_001)
    bump[1:-1] = 1.0                                       # one 899 999-sample plateau above both ends
    assert np.array_equal(ops.find_peaks(bump), scipy_find_peaks(bump)[0])


def test_digital_silence_recording(fe, ref_params):
    """An all-zero recording (and one that is silent for its first 40 s): the envelope is a flat
    run, the reference finds no troughs / peaks there; the GPU path must agree and must not stall."""
    from oracle import ref_port
    sr = 44100
    for pcm in (np.zeros(sr * 60, dtype=np.int16),
                np.concatenate([np.zeros(sr * 40, dtype=np.int16),
                                (3000 * np.sin(2 * np.pi * 50 * np.arange(sr * 20) / sr) *
                                 (np.arange(sr * 20) % (sr // 2) < 2000)).astype(np.int16)])):
        o = ref_port.front_end(pcm, sr, ref_params)
        env, rate, _, _ = fe.preprocess_pcm(pcm, sr, ref_params)
        assert rel_err(env, o["envelope"]) < TOL
        floor, tr = fe._calculate_dynamic_noise_floor(o["envelope"], rate, ref_params)
        assert np.array_equal(tr, o["troughs"])
        assert rel_err(floor.values, o["floor"]) < TOL

        class Clf:
            pass
        c = Clf()
        c.audio_envelope, c.sample_rate, c.params = o["envelope"], rate, ref_params
        assert np.array_equal(fe._find_raw_peaks(c, o["floor"]), o["peaks"])


def test_find_peaks_height_array(ops):
    rng = np.random.default_rng(9)
    x = np.abs(rng.standard_normal(30000))
    h = np.abs(rng.standard_normal(30000)) * 0.8
    ref, _ = scipy_find_peaks(x, height=h, prominence=0.4, distance=15)
    assert np.array_equal(ops.find_peaks(x, h, 0.4, 15, 1), ref)


# ----------------------------------------------------------------------------- K5+K6
@pytest.mark.parametrize("m,window,q,gap", [(5000, 3020, 0.2, 130), (20000, 3010, 0.2, 60), (20000, 333, 0.35, 25),
                                            (3000, 5, 0.2, 40), (8000, 4000, 0.5, 300), (4000, 3, 0.2, 50)])
def test_rolling_floor_matches_pandas(m, window, q, gap, ops):
    rng = np.random.default_rng(m + window)
    env = np.abs(rng.standard_normal(m)) + 0.1
    knots = np.unique(np.sort(rng.integers(gap // 2, m - gap // 3, max(m // gap, 5))))
    env[knots[::7]] = env[knots[0]]                    # exact ties between knots
    if len(knots) > 12:
        env[knots[10]] = env[knots[11]]                # a flat segment
    s = pd.Series(index=knots, data=env[knots]).reindex(np.arange(m)).interpolate()
    ref = s.rolling(window=window, min_periods=3, center=True).quantile(q).bfill().ffill().values
    got = ops.rolling_floor(env, knots, window, q)
    assert rel_err(got, ref) < TOL


@pytest.mark.parametrize("m,window,q,burst", [(400000, 3010, 0.2, False), (400000, 3010, 0.2, True),
                                              (300000, 6660, 0.3, True), (250000, 1505, 0.1, True),
                                              (120000, 13000, 0.2, True)])
def test_rolling_floor_bit_exact_multi_tile(m, window, q, burst, ops):
    """Several CTAs per recording, heavy-tailed knot values and (burst=True) level jumps inside
    a tile, which defeat the pivot estimate and force the sort-everything repeat.  The result
    must be pandas' bit for bit either way."""
    rng = np.random.default_rng(m + window)
    level = np.ones(m)
    if burst:
        for _ in range(12):
            s0 = int(rng.integers(0, m - 5000))
            level[s0:s0 + int(rng.integers(200, 9000))] *= float(rng.choice([0.05, 8.0, 40.0]))
    env = (np.abs(rng.standard_normal(m)) ** 3 + 0.05) * level
    knots = np.unique(np.sort(rng.integers(40, m - 40, m // 85)))
    s = pd.Series(index=knots, data=env[knots]).reindex(np.arange(m)).interpolate()
    ref = s.rolling(window=window, min_periods=3, center=True).quantile(q).bfill().ffill().values
    got = ops.rolling_floor(env, knots, window, q)
    assert np.array_equal(got, ref)


# ----------------------------------------------------------------------------- a2, a3, a4 on reference envelopes
@pytest.mark.parametrize("name", ["vulpine"] + ["synth_" + s for s in SYNTH])
def test_noise_floor_and_peaks_on_reference_envelope(name, fe, ref_params):
    g = load_golden(name)
    env, rate = g["envelope"], int(g["rate"])
    floor, troughs = fe._calculate_dynamic_noise_floor(env, rate, ref_params)
    assert isinstance(floor, pd.Series) and floor.index.dtype == np.int64 and len(floor) == len(env)
    assert np.array_equal(troughs, g["troughs"])                       # bit-exact indices
    assert rel_err(floor.values, g["floor"]) < TOL

    class Clf:
        pass
    c = Clf()
    c.audio_envelope, c.sample_rate, c.params = env, rate, ref_params
    ref_floor = pd.Series(g["floor"], index=np.arange(len(env)))
    peaks = fe._find_raw_peaks(c, ref_floor.values)                    # fed the reference's own floor
    assert peaks.dtype == np.int64 and np.array_equal(peaks, g["raw_peaks"])
    st = fe._initialize_state(c, None, ref_floor, g["troughs"])
    assert np.array_equal(st["all_peaks"], g["raw_peaks"])
    assert st["long_term_bpm"] == 80.0 and st["sorted_troughs"] == sorted(g["troughs"])
    if len(g["raw_peaks"]) >= 2:
        sd = st["smoothed_dev_series"]
        assert np.array_equal(sd.index.values, g["smoothed_dev_index"])
        assert rel_err(sd.values, g["smoothed_dev_values"]) < TOL
    # end to end on the GPU's own floor: expected identical
    assert np.array_equal(fe._find_raw_peaks(c, floor.values), g["raw_peaks"])


def _two_survivor_envelope(inner):
    m = 6000
    tpos = np.array([100, 2500, 2900, 3100, 3500, 5900])
    tval = np.array([1.0, inner, inner + 1, inner - 1, inner + .5, 1.2])
    env = np.interp(np.arange(m), tpos, tval)
    for a, b in zip(tpos[:-1], tpos[1:]):
        x = np.arange(a, b + 1)
        env[a:b + 1] += 60 * np.sin(np.pi * (x - a) / (b - a)) ** 2
    env[:100] += np.linspace(30, 0, 100)
    env[5900:] += np.linspace(0, 30, 100)
    return env


@pytest.mark.parametrize("inner,rate", [(400.0, 600), (200.0, 600), (400.0, 100), (400.0, 40)])
def test_noise_floor_draft_fallback_branches(inner, rate, fe, ref_params):
    """bpm_analysis.py:1107-1110: when <= 2 troughs survive sanitisation the final floor IS the
    draft floor (here: recomputed over all troughs).  (400, 600) leaves exactly two survivors;
    the other cases walk the neighbouring branches (4 kept; local windows; windows of a few
    samples handled by the per-thread kernel)."""
    from oracle import ref_port
    env = _two_survivor_envelope(inner)
    o_floor, o_tr = ref_port.calculate_dynamic_noise_floor(env, rate, ref_params)
    if (inner, rate) == (400.0, 600):
        assert len(o_tr) == 2
    floor, tr = fe._calculate_dynamic_noise_floor(env, rate, ref_params)
    assert np.array_equal(tr, o_tr)
    assert np.array_equal(floor.values, o_floor.values)


def test_vulpine_matches_shipped_debug_log(fe, ref_params):
    """Raw-peak indices the reference itself logged for its sample recording."""
    g = load_golden("vulpine")

    class Clf:
        pass
    c = Clf()
    c.audio_envelope, c.sample_rate, c.params = g["envelope"], int(g["rate"]), ref_params
    floor, troughs = fe._calculate_dynamic_noise_floor(g["envelope"], int(g["rate"]), ref_params)
    assert np.array_equal(fe._find_raw_peaks(c, floor.values), g["log_peak_idx"])
    assert np.sum(troughs != g["log_trough_idx"]) <= 4


# ----------------------------------------------------------------------------- a5..a8
@pytest.mark.parametrize("name", ["vulpine", "synth_c2_240s", "synth_holter_180s", "synth_c1_30s"])
def test_beat_reductions_match_reference(name, fe, ref_params):
    g = load_golden(name)
    beats, rate = g["beats"], int(g["rate"])
    series, times = fe.calculate_bpm_series(beats, rate, ref_params)
    assert np.array_equal(series.index.as_unit("us").asi8, g["bpm_index_us"])
    assert str(series.index.dtype) == str(g["bpm_index_unit"])
    assert np.array_equal(times, g["bpm_times"])
    assert rel_err(series.values, g["bpm_values"]) < TOL
    # downstream reductions are fed the reference's own series (bit-exact decisions)
    ref_series = pd.Series(g["bpm_values"], index=pd.DatetimeIndex(g["bpm_index_us"].astype("datetime64[us]")))
    for key, fn in (("recovery", fe.find_peak_recovery_rate), ("exertion", fe.find_peak_exertion_rate)):
        d = fn(ref_series)
        assert (d is not None) == bool(int(g[key + "_present"]))
        if d is not None:
            assert pd.Timestamp(d["start_time"]).as_unit("us").value == int(g[key + "_start_us"])
            assert pd.Timestamp(d["end_time"]).as_unit("us").value == int(g[key + "_end_us"])
            assert d["slope_bpm_per_sec"] == float(g[key + "_slope_bpm_per_sec"])
            assert d["duration_sec"] == float(g[key + "_duration_sec"])
    inc, dec = fe.find_major_hr_inclines(ref_series), fe.find_major_hr_declines(ref_series)
    assert len(inc) == int(g["inclines_n"]) and len(dec) == int(g["declines_n"])
    assert np.array_equal([pd.Timestamp(d["start_time"]).as_unit("us").value for d in inc], g["inclines_start_us"])
    assert np.array_equal([d["slope_bpm_per_sec"] for d in dec], g["declines_slope_bpm_per_sec"])
    hrv = fe.calculate_windowed_hrv(beats, rate, ref_params)
    assert list(hrv.columns) == ["time", "rmssdc", "sdnn", "bpm"]
    assert len(hrv) == len(g["hrv"])
    if len(hrv):
        assert rel_err(hrv.values.astype(np.float64), g["hrv"]) < TOL


def test_beat_reduction_edge_cases(fe, ref_params):
    s, t = fe.calculate_bpm_series(np.array([5]), 302, ref_params)
    assert s.empty and len(t) == 0
    s, t = fe.calculate_bpm_series(np.array([5, 5, 5]), 302, ref_params)      # no interval > 1e-6
    assert s.empty and len(t) == 0
    assert fe.calculate_windowed_hrv(np.arange(0, 39 * 300, 300), 302, ref_params).empty
    assert fe.calculate_windowed_hrv(np.arange(0, 40 * 300, 300), 302, ref_params).empty   # 39 intervals
    assert len(fe.calculate_windowed_hrv(np.arange(0, 41 * 300, 300), 302, ref_params)) == 1
    assert fe.find_peak_exertion_rate(pd.Series(dtype=np.float64)) is None


# ----------------------------------------------------------------------------- batch, one C call
def test_stage_a_ragged_batch_matches_oracle(fe, ref_params, synth_inputs):
    import torch
    from bpm_analysis_b200.runtime import StageARunner
    names = ["c1_30s", "i32_20s", "c1_30s"]
    # same dtype within a batch: int16 items of different lengths
    pcms = [synth_inputs["c1_30s"][0], synth_inputs["c1_30s"][0][: 44100 * 11], synth_inputs["stereo_20s"][0][:, 0].copy()]
    sr = 44100
    runner = StageARunner([len(p) for p in pcms], sr, ref_params, np.int16, 1, want_debug=False)
    runner.upload(pcms)
    runner.launch()
    torch.cuda.synchronize()
    res = runner.result()
    from oracle import ref_port
    for i, pcm in enumerate(pcms):
        o = ref_port.front_end(pcm, sr, ref_params)
        r = res.item(i)
        assert rel_err(r["envelope"], o["envelope"]) < TOL
        assert np.array_equal(r["troughs"], o["troughs"])
        assert np.array_equal(r["peaks"], o["peaks"])
        assert rel_err(r["floor"], o["floor"]) < TOL
        assert rel_err(r["smoothed_dev"], o["smoothed_dev_series"].values) < TOL
    del names


def test_full_size_properties_c2(fe, ref_params):
    """BASELINE configs[1] size (60 min @ 48 kHz), checked through size-independent properties:
    linearity of the filter, envelope >= 0, floor within the envelope's range, sorted unique
    indices, peaks above floor, idempotent re-run."""
    import torch
    from bpm_analysis_b200 import synth
    from bpm_analysis_b200.runtime import StageARunner
    pcm, sr, _ = synth.config_c2(seed=2, duration_sec=3600.0)
    runner = StageARunner([len(pcm)], sr, ref_params)
    runner.upload([pcm])
    runner.launch()
    torch.cuda.synchronize()
    a = {k: v.clone() for k, v in runner.out.items()}
    r = runner.result().item(0)
    env, floor, peaks, troughs = r["envelope"], r["floor"], r["peaks"], r["troughs"]
    assert len(env) == 1086793 and np.all(env >= 0) and np.all(np.isfinite(env))
    assert np.all(np.isfinite(floor)) and floor.min() >= env.min() and floor.max() <= env.max()
    for idx in (peaks, troughs):
        assert np.all(np.diff(idx) >= 15) and idx.min() > 0 and idx.max() < len(env) - 1
    assert np.all(env[peaks] >= floor[peaks])
    # ~2 raw peaks per beat over the 60->170->80 BPM ramp
    assert 9000 < len(peaks) < 20000
    # linearity: filtering 2x the input doubles the band-passed signal (int16 headroom kept)
    half = (pcm // 2).astype(np.int16)
    runner.upload([(half * 2).astype(np.int16)])
    runner.launch()
    torch.cuda.synchronize()
    f2 = runner.out["filtered"].clone()
    runner.upload([half])
    runner.launch()
    torch.cuda.synchronize()
    f1 = runner.out["filtered"]
    assert float((f2 - 2.0 * f1).abs().max() / f2.abs().max()) < 1e-12
    # idempotence of the whole stage
    runner.upload([pcm])
    runner.launch()
    torch.cuda.synchronize()
    for k in ("envelope", "floor"):
        assert torch.equal(runner.out[k], a[k])
    n = int(runner.out["peak_count"][0])
    assert torch.equal(runner.out["peaks"][:n], a["peaks"][:n])


# ----------------------------------------------------------------------------- K8b (parity unpinned)
@pytest.mark.parametrize("name", ["vulpine", "synth_c2_240s", "synth_holter_180s"])
def test_peak_trough_noise_matches_oracle(name, fe, ref_params):
    """Optional surrounding-trough noise metric: no reference function exists (SURVEY §8a note), so
    the check is against the oracle's restatement of the documented rule -- bit for bit."""
    from oracle import ref_port
    g = load_golden(name)
    env, floor, peaks, troughs = g["envelope"], g["floor"], g["raw_peaks"], g["troughs"]
    p = dict(ref_params, trough_noise_multiplier=1.5)              # both flag values occur
    got = fe.peak_trough_noise(env, floor, peaks, troughs, p)
    ref = ref_port.peak_trough_noise(env, floor, peaks, troughs, p)
    for k in ("prev_amp", "next_amp", "ratio"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    assert np.array_equal(got["flags"], ref["flags"])
    assert got["flags"].max() >= 1
    # no troughs at all: everything NaN, no flags
    none = fe.peak_trough_noise(env, floor, peaks, np.array([], dtype=np.int64), p)
    assert np.all(np.isnan(none["ratio"])) and not none["flags"].any()


# ----------------------------------------------------------------------------- ingest pipeline
@pytest.mark.parametrize("ingest", ["ce", "sm"])
def test_pipelined_zero_copy_ingest_is_bit_identical(fe, ref_params, ingest):
    """runtime.StageAPipeline (kept frames out of pinned host memory by bpm_copy_frames on the
    copy engine / by bpm_gather_frames from the SMs, compute and read-back on three streams, two
    recordings in flight) must return exactly what the one-shot StageARunner returns for each
    recording of a stream of equal-length recordings."""
    import torch
    from bpm_analysis_b200 import synth
    from bpm_analysis_b200.runtime import StageAPipeline, StageARunner
    recs = [synth.config_c2(seed=20 + i, duration_sec=200.0)[0] for i in range(5)]
    sr = 48000
    one = StageARunner([len(recs[0])], sr, ref_params)
    want = []
    for r in recs:
        one.upload([r])
        one.launch()
        torch.cuda.synchronize()
        want.append({k: v.clone().cpu() for k, v in one.out.items()})
    for use_graph in (False, True):
        pipe = StageAPipeline(len(recs[0]), sr, ref_params, depth=2, use_graph=use_graph, ingest=ingest)
        pins = [torch.from_numpy(r).pin_memory() for r in recs]
        got = []
        for k in range(len(recs)):
            if k >= 2:
                got.append({n: t.clone() for n, t in pipe.wait(k - 2).items()})
            pipe.submit(k, pins[k])
        for k in range(len(recs) - 2, len(recs)):
            got.append({n: t.clone() for n, t in pipe.wait(k).items()})
        for g, w in zip(got, want):
            nt, npk = int(w["trough_count"][0]), int(w["peak_count"][0])
            assert int(g["trough_count"][0]) == nt and int(g["peak_count"][0]) == npk
            assert torch.equal(g["envelope"], w["envelope"]) and torch.equal(g["floor"], w["floor"])
            assert torch.equal(g["troughs"][:nt], w["troughs"][:nt]) and torch.equal(g["peaks"][:npk], w["peaks"][:npk])
            assert torch.equal(g["smoothed_dev"][:npk - 1], w["smoothed_dev"][:npk - 1])


@pytest.mark.parametrize("kind", ["stereo_i16", "f32", "u8", "stereo_f64"])
def test_copy_engine_ingest_other_formats(fe, ref_params, kind):
    """bpm_copy_frames keeps the PCM's dtype and channels: a pregathered='ce' runner fed from pinned
    host memory equals the fused one-shot runner bit for bit for multi-channel and non-int16 input."""
    import torch
    from bpm_analysis_b200 import synth
    from bpm_analysis_b200.runtime import StageARunner
    pcm, sr, _ = synth.config_c1(seed=31, duration_sec=47.3)
    if kind == "stereo_i16":
        pcm = np.stack([pcm, np.roll(pcm, 7) // 2], axis=1)
    elif kind == "f32":
        pcm = (pcm / 32768.0).astype(np.float32)
    elif kind == "u8":
        pcm = ((pcm.astype(np.int32) >> 8) + 128).astype(np.uint8)
    elif kind == "stereo_f64":
        pcm = np.stack([pcm / 3.0, np.roll(pcm, 3) / 7.0], axis=1)
    channels = 1 if pcm.ndim == 1 else pcm.shape[1]
    one = StageARunner([pcm.shape[0]], sr, ref_params, pcm.dtype, channels)
    one.upload([pcm])
    one.launch()
    ce = StageARunner([pcm.shape[0]], sr, ref_params, pcm.dtype, channels, pregathered="ce")
    pin = torch.from_numpy(np.ascontiguousarray(pcm).reshape(-1)).pin_memory()
    ce.gather(pin)
    ce.launch()
    torch.cuda.synchronize()
    nt, npk = int(one.out["trough_count"][0]), int(one.out["peak_count"][0])
    assert nt > 20 and npk > 20
    assert int(ce.out["trough_count"][0]) == nt and int(ce.out["peak_count"][0]) == npk
    for k in ("filtered", "envelope", "floor"):
        assert torch.equal(ce.out[k], one.out[k]), k
    assert torch.equal(ce.out["troughs"][:nt], one.out["troughs"][:nt])
    assert torch.equal(ce.out["peaks"][:npk], one.out["peaks"][:npk])


# ----------------------------------------------------------------------------- randomized sweep
@pytest.mark.timeout(300)
@pytest.mark.parametrize("seed", list(range(16)))
def test_randomized_recordings_match_oracle(seed, fe, ref_params):
    """Random sample rates, durations, channel counts, sample formats, amplitudes and parameter
    values (seeded): whatever the oracle returns -- or raises -- for a1..a4, the GPU path must too."""
    import pandas as pd
    from bpm_analysis_b200 import synth
    from oracle import ref_port
    rng = np.random.default_rng(1000 + seed)
    sr = int(rng.choice([4000, 8000, 11025, 16000, 22050, 44100, 48000, 96000]))
    dur = float(rng.uniform(0.4, 25.0))
    bpm = float(rng.uniform(45, 180))
    pcm, _, _ = synth.pcg_recording(dur, sr, lambda t: bpm, seed=seed)
    kind = rng.choice(["i16", "i16", "f32", "u8", "i32", "f64", "stereo"])
    if kind == "f32":
        pcm = (pcm / 32768.0).astype(np.float32)
    elif kind == "f64":
        pcm = pcm / 32768.0 * float(rng.uniform(1e-3, 1e3))
    elif kind == "u8":
        pcm = ((pcm.astype(np.int32) >> 8) + 128).astype(np.uint8)
    elif kind == "i32":
        pcm = pcm.astype(np.int32) * 65536
    elif kind == "stereo":
        pcm = np.stack([pcm, (pcm * 0.5).astype(np.int16)], axis=1)
    p = dict(ref_params)
    p["downsample_factor"] = int(rng.choice([1, 7, 50, 300, 1000]))
    p["noise_window_sec"] = float(rng.choice([0.5, 2, 10, 30]))
    p["noise_floor_quantile"] = float(rng.choice([0.05, 0.2, 0.5]))
    p["min_peak_distance_sec"] = float(rng.choice([0.02, 0.05, 0.2]))
    p["peak_prominence_quantile"] = float(rng.choice([0.1, 0.1, 0.3]))
    p["trough_rejection_multiplier"] = float(rng.choice([1.5, 4.0]))
    try:
        o = ref_port.front_end(pcm, sr, p)
    except Exception as e:                                   # noqa: BLE001 - the reference's own failure modes
        with pytest.raises(type(e)):
            env, rate, _, _ = fe.preprocess_pcm(pcm, sr, p)
            floor, tr = fe._calculate_dynamic_noise_floor(env, rate, p)
            class C0:
                pass
            c0 = C0()
            c0.audio_envelope, c0.sample_rate, c0.params = env, rate, p
            fe._find_raw_peaks(c0, floor.values)
        return
    env, rate, filt, _ = fe.preprocess_pcm(pcm, sr, p)
    assert rate == o["rate"]
    # The reference filters with the transfer-function form (butter(...)->(b, a), filtfilt).  At the
    # default decimated rate (~300 Hz) it agrees with the exact (SOS / state-space) filter to 6e-14; with
    # decimation switched off the 20 Hz edge sits at 1e-3 of Nyquist and scipy's OWN tf-form result is
    # 7e-11 (11 kHz) ... 7e-8 (44.1 kHz) away from its sosfiltfilt: that rounding noise of an
    # ill-conditioned recursion is not reproducible by a reordered evaluation, so the envelope tolerance
    # follows it there.  Index parity below is unaffected (it is asserted on the reference's envelope).
    assert rel_err(env, o["envelope"]) < (TOL if rate <= 12000 else 1e-6)
    # index outputs on the reference's own envelope, floats to 1e-9
    floor, tr = fe._calculate_dynamic_noise_floor(o["envelope"], rate, p)
    assert np.array_equal(tr, o["troughs"])
    assert rel_err(floor.values, o["floor"]) < TOL

    class Clf:
        pass
    c = Clf()
    c.audio_envelope, c.sample_rate, c.params = o["envelope"], rate, p
    assert np.array_equal(fe._find_raw_peaks(c, o["floor"]), o["peaks"])
