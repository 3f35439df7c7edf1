"""BASELINE.json configs beyond the bench workload, at (or scaled from) their stated sizes:
C3 batch of 10-min recordings, C4 24-h Holter stream, C5 parameter sweep.  Oracle parity on
what the CPU finishes in seconds, size-independent properties at full size."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    from bpm_analysis_b200 import _native
    _native.load_library()
    return torch


def test_c3_batch_of_10min_recordings(gpu, ref_params):
    """C3: one rank's share is 128 ragged 10-min 44.1 kHz recordings in ONE bpm_stage_a call.
    16 distinct recordings are tiled 8x (with different tails cut off, so offsets differ):
    twins must agree bit for bit wherever they overlap in time, two items are checked against
    the oracle."""
    from bpm_analysis_b200 import synth
    from bpm_analysis_b200.runtime import StageARunner
    from oracle import ref_port
    base = [synth.config_c3_item(i)[0] for i in range(16)]
    sr = 44100
    cuts = [0, 441, 44100 * 3 + 7, 146 * 1000, 5, 44100 * 30, 1, 146 * 777 + 3]
    pcms = [base[i % 16][: len(base[i % 16]) - cuts[i // 16]] for i in range(128)]
    runner = StageARunner([len(p) for p in pcms], sr, ref_params)
    runner.upload(pcms)
    runner.launch()
    gpu.cuda.synchronize()
    res = runner.result()
    assert runner.total_in == sum(len(p) for p in pcms) and runner.n_items == 128
    for i in (0, 37):
        o = ref_port.front_end(pcms[i], sr, ref_params)
        r = res.item(i)
        assert rel_err(r["envelope"], o["envelope"]) < TOL
        assert rel_err(r["floor"], o["floor"]) < TOL
        assert np.array_equal(r["troughs"], o["troughs"]) and np.array_equal(r["peaks"], o["peaks"])
    # the untruncated twins of an item (cut 0 and cut 1 share all but the last decimated sample)
    for i in range(16):
        a, b = res.item(i), res.item(i + 16 * 6)
        k = len(b["envelope"]) - 4000          # away from the end (zero-phase filter: the tail differs)
        assert rel_err(a["envelope"][:k], b["envelope"][:k]) < 1e-12
        assert np.array_equal(a["peaks"][a["peaks"] < k - 4000], b["peaks"][b["peaks"] < k - 4000])
    counts = res.host("peak_count")
    assert np.all(counts > 600) and np.all(counts < 4000)          # ~2 raw peaks per beat, 55..95 BPM


def test_c4_24h_holter_stream_full_size(gpu, ref_params):
    """C4: 24 h at 4 kHz (M = 28.8 M envelope samples), bursts and dropouts.  Properties at
    full size + the chunked (4 ranks) run must reproduce the one-shot run."""
    from bpm_analysis_b200 import stream, synth
    from bpm_analysis_b200.runtime import StageARunner
    pcm, sr, beats = synth.config_c4(seed=4, duration_sec=86400.0)
    runner = StageARunner([len(pcm)], sr, ref_params)
    runner.upload([pcm])
    runner.launch()
    gpu.cuda.synchronize()
    r = runner.result().item(0)
    env, floor, peaks, troughs = r["envelope"], r["floor"], r["peaks"], r["troughs"]
    assert len(env) == 28800000 and np.all(env >= 0) and np.all(np.isfinite(env))
    assert np.all(np.isfinite(floor)) and floor.min() >= env.min() and floor.max() <= env.max()
    d = int(ref_params["min_peak_distance_sec"] * r["rate"])
    for idx in (peaks, troughs):
        assert np.all(np.diff(idx) >= d) and idx.min() > 0 and idx.max() < len(env) - 1
    assert np.all(env[peaks] >= floor[peaks])
    assert 1.5 * len(beats) < len(peaks) < 3.0 * len(beats)
    # every kept trough passes the reference's rejection rule against the FINAL floor's draft
    # only approximately; what must hold exactly is idempotence:
    runner.launch()
    gpu.cuda.synchronize()
    r2 = runner.result().item(0)
    assert np.array_equal(r2["floor"], floor) and np.array_equal(r2["peaks"], peaks)
    # chunked over 4 ranks (threads sharing the GPU): same lists, same floats to 1e-9
    eng = stream.DeviceEngine()

    def body(comm):
        fe = stream.ChunkedFrontEnd(len(pcm), sr, ref_params, comm, eng)
        f0, f1 = fe.frames()
        out = fe.run(eng.tensor(pcm[f0:f1]))
        gpu.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items() if k in ("envelope", "floor", "troughs", "peaks")}

    for got in stream.run_thread_world(4, body)[:2]:
        assert rel_err(got["envelope"], env) < TOL and rel_err(got["floor"], floor) < TOL
        assert np.array_equal(got["troughs"], troughs) and np.array_equal(got["peaks"], peaks)


def test_c5_parameter_sweep(gpu, ref_params):
    """C5: 256 settings (16 band-passes x 16 noise-floor settings) over one recording.  A
    5-min recording is checked against the oracle on 10 settings; the 30-min recording runs
    all 256, sharded over 2 'ranks' whose blocks must tile the sweep."""
    from bpm_analysis_b200 import sweep, synth
    from oracle import ref_port
    settings = synth.c5_settings()
    assert len(settings) == 256
    pcm, sr, _ = synth.config_c5(seed=5, duration_sec=300.0)
    pick = [0, 17, 63, 64, 100, 150, 201, 202, 240, 255]
    got = sweep.run_sweep(pcm, sr, ref_params, [settings[i] for i in pick], keep_arrays=True)
    for rec, i in zip(got, pick):
        p = dict(ref_params, **settings[i])
        if "error" in rec:                       # 100 Hz upper edge: decimated Nyquist == 100 Hz (:1041)
            with pytest.raises(ValueError):
                ref_port.front_end(pcm, sr, p)
            assert settings[i]["highcut_hz"] == 100.0
            continue
        o = ref_port.front_end(pcm, sr, p)
        assert rec["rate"] == o["rate"]
        assert rel_err(rec["envelope"].cpu().numpy(), o["envelope"]) < TOL
        assert rel_err(rec["floor"].cpu().numpy(), o["floor"]) < TOL
        assert np.array_equal(rec["troughs"].cpu().numpy(), o["troughs"])
        assert np.array_equal(rec["peaks"].cpu().numpy(), o["peaks"])
    pcm, sr, _ = synth.config_c5(seed=5, duration_sec=1800.0)
    parts = [sweep.run_sweep(pcm, sr, ref_params, settings, rank=r, world=2) for r in range(2)]
    flat = parts[0] + parts[1]
    assert [x["setting"] for x in flat] == list(range(256))
    ok = [x for x in flat if "error" not in x]
    assert len(ok) == 192 and all(settings[x["setting"]]["highcut_hz"] == 100.0 for x in flat if "error" in x)
    assert all(x["n_peaks"] > 1000 and x["n_troughs"] > 1000 for x in ok)
    # settings that differ only in the noise floor share the band-pass, hence rate and length
    assert len({(x["rate"], x["m"]) for x in flat[16:32]}) == 1


def _dropin_chain(frontend, pcm, sr, params, beat_idx):
    """The ten drop-in calls in analyze_wav_file's order on host arrays (bpm_analysis.py:1731-1757)."""
    class _Clf:
        pass
    clf = _Clf()
    env, rate, _, _ = frontend.preprocess_pcm(pcm, sr, params, want_filtered=False)
    floor, troughs = frontend._calculate_dynamic_noise_floor(env, rate, params)
    clf.audio_envelope, clf.sample_rate, clf.params = env, rate, params
    st1 = frontend._initialize_state(clf, None, floor, troughs)
    st2 = frontend._initialize_state(clf, 80.0, floor, troughs)
    sm, bt = frontend.calculate_bpm_series(beat_idx, rate, params)
    return {"envelope": env, "rate": rate, "floor": floor, "troughs": troughs, "peaks": st2["all_peaks"],
            "peaks_first": st1["all_peaks"], "smoothed_dev": st2["smoothed_dev_series"], "smoothed_bpm": sm,
            "bpm_times": bt, "major_inclines": frontend.find_major_hr_inclines(sm),
            "major_declines": frontend.find_major_hr_declines(sm), "hrr": frontend.calculate_hrr(sm),
            "recovery": frontend.find_peak_recovery_rate(sm), "exertion": frontend.find_peak_exertion_rate(sm),
            "hrv": frontend.calculate_windowed_hrv(beat_idx, rate, params)}


def _check_against_oracle(got, fe, br):
    assert rel_err(got["envelope"], fe["envelope"]) < TOL
    assert rel_err(got["floor"].values, fe["floor"]) < TOL
    assert np.array_equal(got["troughs"], fe["troughs"])
    assert np.array_equal(got["peaks"], fe["peaks"]) and np.array_equal(got["peaks_first"], fe["peaks"])
    assert rel_err(got["smoothed_dev"].values, fe["smoothed_dev_series"].values) < TOL
    assert np.array_equal(got["smoothed_dev"].index.values, fe["smoothed_dev_series"].index.values)
    assert got["smoothed_bpm"].index.equals(br["smoothed_bpm"].index)
    assert rel_err(got["smoothed_bpm"].values, br["smoothed_bpm"].values) < TOL
    assert np.array_equal(got["bpm_times"], br["bpm_times"])
    for a, b in ((got["recovery"], br["peak_recovery_stats"]), (got["exertion"], br["peak_exertion_stats"])):
        assert (a is None) == (b is None)
        if a is not None:
            assert a["start_time"] == b["start_time"] and a["end_time"] == b["end_time"]
            assert a["slope_bpm_per_sec"] == b["slope_bpm_per_sec"] and a["duration_sec"] == b["duration_sec"]
    for a, b in ((got["major_inclines"], br["major_inclines"]), (got["major_declines"], br["major_declines"])):
        assert [(x["start_time"], x["end_time"]) for x in a] == [(x["start_time"], x["end_time"]) for x in b]
    assert rel_err(got["hrv"].values, br["windowed_hrv_df"].values) < TOL


def test_full_size_c2_matches_oracle_through_the_drop_in_functions(gpu, ref_params):
    """The bench configuration itself (BASELINE configs[1]: 60 min @ 48 kHz, M = 1 086 793) against
    the CPU oracle, through the drop-in functions: envelope / floor / deviation series to 1e-9,
    trough and raw-peak lists and every beat-list reduction exact -- on the GPU's own envelope."""
    from bpm_analysis_b200 import dropin, frontend, synth
    from oracle import ref_port
    pcm, sr, beats = synth.config_c2(seed=2)
    d = dropin.dropin()
    d.forget()
    before = dict(d.stats)
    fe = ref_port.front_end(pcm, sr, ref_params)
    beat_idx = synth.beats_to_envelope_indices(beats, fe["rate"])
    br = ref_port.beat_reductions(beat_idx, fe["rate"], ref_params)
    got = _dropin_chain(frontend, pcm, sr, ref_params, beat_idx)
    assert len(got["envelope"]) == 1086793 and got["rate"] == 301
    _check_against_oracle(got, fe, br)
    assert len(fe["peaks"]) > 10000 and len(br["windowed_hrv_df"]) > 1000
    after = dict(d.stats)
    assert after["stage_a_calls"] - before["stage_a_calls"] == 1        # the whole chain cost one stage-A call
    assert after["session_misses"] == before["session_misses"]          # nothing was uploaded a second time


def test_c4_two_hours_match_oracle_one_shot_and_chunked(gpu, ref_params):
    """C4 (4 kHz Holter stream with bursts and dropouts), first 2 h (M = 2.4 M): the one-shot drop-in
    chain AND the halo-chunked front end (4 ranks) against the CPU oracle."""
    from bpm_analysis_b200 import frontend, stream, synth
    from oracle import ref_port
    pcm, sr, beats = synth.config_c4(seed=4, duration_sec=7200.0)
    fe = ref_port.front_end(pcm, sr, ref_params)
    beat_idx = synth.beats_to_envelope_indices(beats, fe["rate"])
    br = ref_port.beat_reductions(beat_idx, fe["rate"], ref_params)
    got = _dropin_chain(frontend, pcm, sr, ref_params, beat_idx)
    assert len(got["envelope"]) == 2400000
    _check_against_oracle(got, fe, br)
    eng = stream.DeviceEngine()

    def body(comm):
        f = stream.ChunkedFrontEnd(len(pcm), sr, ref_params, comm, eng)
        f0, f1 = f.frames()
        out = f.run(eng.tensor(pcm[f0:f1]))
        gpu.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items() if k in ("envelope", "floor", "troughs", "peaks")}

    for ch in stream.run_thread_world(4, body)[:2]:
        assert rel_err(ch["envelope"], fe["envelope"]) < TOL and rel_err(ch["floor"], fe["floor"]) < TOL
        assert np.array_equal(ch["troughs"], fe["troughs"]) and np.array_equal(ch["peaks"], fe["peaks"])
