"""The compiled sequential classifier (libbpm_host.so, include/bpm_host.h) against the reference's
``PeakClassifier.classify_peaks`` (bpm_analysis.py:113-329, :1120-1258): candidate beats, every
per-peak debug string, the long-term BPM trace and the log lines must be IDENTICAL (bit for bit,
byte for byte).  Golden outputs of the unmodified reference: tests/golden/classifier_cases.json.gz
(oracle/make_golden_classifier.py); where the reference tree is present the comparison is also
made live on many more seeds and on full recordings through the reference's own constructor.
"""
import collections
import ctypes
import gzip
import json
import logging
import os

import numpy as np
import pytest

from _classifier_cases import ClassifierStandIn, make_case, pack_result
from bpm_analysis_b200 import classifier
from conftest import GOLDEN_DIR, REPO
from oracle.load_reference import load_reference, reference_available

needs_reference = pytest.mark.skipif(not reference_available(), reason="needs /root/reference")


def run_ours(case, caplog):
    obj = ClassifierStandIn(case)
    caplog.clear()
    with caplog.at_level(logging.INFO):
        packed = pack_result(classifier.classify_peaks(obj))
    packed["log"] = [r.getMessage() for r in caplog.records]
    packed["final_long_term_bpm"] = float(obj.state["long_term_bpm"]).hex()
    packed["final_consecutive_rr_rejections"] = int(obj.state["consecutive_rr_rejections"])
    return packed, obj


def assert_same(got, want, label):
    assert got["final_peaks"] == want["final_peaks"], label
    assert got["keys"] == want["keys"], label
    for k, a, b in zip(want["keys"], got["texts"], want["texts"]):
        assert a == b, f"{label}: debug string of peak {k}"
    assert got["lt_times"] == want["lt_times"], label
    assert got["lt_values"] == want["lt_values"], label
    assert got["log"] == want["log"], label
    assert got["final_long_term_bpm"] == want["final_long_term_bpm"], label
    assert got["final_consecutive_rr_rejections"] == want["final_consecutive_rr_rejections"], label


def test_host_library_exports_header_symbols():
    import re
    lib = ctypes.CDLL(classifier.HOST_LIB_PATH)
    header = open(os.path.join(REPO, "include", "bpm_host.h")).read()
    from bpm_analysis_b200 import corrections
    declared = set(re.findall(r"^\s*(?:int|void|int64_t)\s+(bpm_\w+)\s*\(", header, flags=re.M))
    assert declared == set(classifier.EXPORTED_SYMBOLS) | set(corrections.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert classifier.load_host_library().bpm_host_abi_version() == classifier.HOST_ABI_VERSION
    assert ctypes.sizeof(classifier.ClassifierParams) == 28 * 8 + 4 * 4
    assert ctypes.sizeof(classifier.ClassifyJob) == 10 * 8


def test_fixed_point_formatter_matches_python():
    """Every number in a debug string goes through this formatter; it must print what Python's
    format(v, '.Nf') prints, including exact ties (round-half-even on the binary value), values
    just off a tie, negative zero results, huge values, inf and nan."""
    lib = classifier.load_host_library()
    buf = ctypes.create_string_buffer(512)
    rng = np.random.default_rng(11)
    values = [0.0, -0.0, 0.125, 0.375, 2.5, 3.5, 0.5, 1.5, -0.5, -0.001, 0.005, 0.015, 0.025, 1.005, 2.675, 1e-9,
              0.9999999, 0.99999, 999.9995, 1234567.8915, 2.0 ** 40, 2.0 ** 40 - 0.5, 1e15 + 0.5, 1e22, 1e300,
              float("inf"), float("-inf"), float("nan"), 0.285, 0.28500000000000003, 0.145, 100.0 * 0.125]
    values += list(rng.uniform(0, 2, 4000)) + list(rng.uniform(-300, 300, 2000)) + list(10.0 ** rng.uniform(-12, 14, 2000))
    values += [k / 8.0 for k in range(-40, 200)] + [k / 1000.0 + 0.0005 for k in range(300)]       # ties and near-ties
    values += [np.nextafter(k / 200.0 + 0.0025, s) for k in range(200) for s in (-1e9, 1e9)]
    for prec in (0, 1, 2, 3):
        for v in values:
            n = lib.bpm_host_format_fixed(float(v), prec, buf, len(buf))
            assert n >= 0
            assert buf.value.decode() == format(float(v), f".{prec}f"), (v, prec)


def test_golden_cases_reproduced_exactly(caplog):
    with gzip.open(os.path.join(GOLDEN_DIR, "classifier_cases.json.gz"), "rb") as fh:
        golden = json.loads(fh.read().decode("utf-8"))
    labels = collections.Counter()
    events = 0
    for seed, want in sorted(golden.items(), key=lambda kv: int(kv[0])):
        got, _ = run_ours(make_case(int(seed)), caplog)
        assert_same(got, want, f"seed {seed}")
        labels.update(t.split("§")[0] for t in want["texts"])
        events += len(want["log"])
    # the fixture exercises every label the classifier can assign and both log lines
    assert set(labels) == set(classifier.PEAK_TYPE_LABELS.values()), labels
    assert min(labels.values()) >= 5 and labels["Noise"] >= 100 and labels["Lone S1 (Corrected by Cascade Reset)"] >= 50, labels
    assert events > 10


@needs_reference
def test_live_reference_many_seeds(caplog):
    from oracle.make_golden_classifier import run_reference
    ref = load_reference()
    for seed in range(100, 260):
        case = make_case(seed)
        want = run_reference(ref, case)
        got, _ = run_ours(case, caplog)
        assert_same(got, want, f"seed {seed}")


@needs_reference
def test_full_recording_through_reference_constructor(caplog, tmp_path):
    """Recordings through the reference's own front end and constructor (CPU), then both loops."""
    from scipy.io import wavfile
    from bpm_analysis_b200 import synth
    from oracle.load_reference import reference_params
    ref = load_reference()
    params = reference_params()
    params["save_filtered_wav"] = False
    for seed, (rate, sigma, bpm) in enumerate([(8000, 0.02, lambda t: 70.0), (8000, 0.2, lambda t: 140.0 + 50 * np.sin(t / 15))]):
        pcm, sr, _ = synth.pcg_recording(120.0, rate, bpm, 40 + seed, noise_sigma=sigma)
        path = str(tmp_path / f"r{seed}.wav")
        wavfile.write(path, sr, pcm)
        env, nsr = ref.preprocess_audio(path, params, str(tmp_path))
        floor, troughs = ref._calculate_dynamic_noise_floor(env, nsr, params)
        for thr, hint, window in ((0.75, None, (None, None)), (0.5, 95.0, (30.0, 80.0))):
            p = dict(params, pairing_confidence_threshold=thr)
            a = ref.PeakClassifier(env, nsr, p, hint, floor, troughs, *window)
            b = ref.PeakClassifier(env, nsr, p, hint, floor, troughs, *window)
            want = pack_result(a.classify_peaks())
            got = pack_result(classifier.classify_peaks(b))
            for k in ("final_peaks", "keys", "texts", "lt_times", "lt_values"):
                assert got[k] == want[k], (seed, thr, k)
            assert a.state["long_term_bpm"] == b.state["long_term_bpm"]
            assert [int(x) for x in a.state["candidate_beats"]] == [int(x) for x in b.state["candidate_beats"]]


def test_fewer_than_two_peaks_and_second_call(caplog):
    case = make_case(3)
    one = dict(case, peaks=case["peaks"][:1], dev_index=case["dev_index"][:0], dev_values=case["dev_values"][:0])
    obj = ClassifierStandIn(one)
    final, allp, data = classifier.classify_peaks(obj)                  # bpm_analysis.py:115-116
    assert final is obj.state["all_peaks"] and allp is final and data == {"beat_debug_info": {}}
    obj = ClassifierStandIn(case)
    first = pack_result(classifier.classify_peaks(obj))
    again = pack_result(classifier.classify_peaks(obj))                 # the loop is over: only finalisation repeats
    assert first == again


def test_argument_errors_are_reported():
    case = make_case(2)
    packed = classifier.pack_params(case["params"], 80.0, None, None)
    bad = case["peaks"].copy()
    bad[3] = bad[2]                                                     # not strictly ascending
    with pytest.raises(ValueError):
        classifier.classify_arrays(case["env"], case["floor"], bad, case["dev_index"], case["dev_values"], 300, packed)
    with pytest.raises(ValueError):
        classifier.classify_arrays(case["env"], case["floor"][:-1], case["peaks"], case["dev_index"],
                                   case["dev_values"], 300, packed)
    with pytest.raises(KeyError):                                       # required key, like the reference's params[...]
        classifier.pack_params({k: v for k, v in case["params"].items() if k != "min_bpm"}, 80.0, None, None)


def test_python_level_errors_of_the_first_iteration_are_mirrored():
    case = make_case(4)
    case["params"] = dict(case["params"], contractility_bpm_low=120.0, contractility_bpm_high=120.0)
    with pytest.raises(ZeroDivisionError):                              # bpm_analysis.py:1136, Python floats
        classifier.classify_peaks(ClassifierStandIn(case))
    case = make_case(4)
    case["params"] = dict(case["params"], stability_history_window=0)
    with pytest.raises(ZeroDivisionError):                              # :141, paired_count / 0
        classifier.classify_peaks(ClassifierStandIn(case))


@needs_reference
def test_live_reference_adversarial_inputs(caplog):
    """NaNs in the deviation series and in the floor, amplitudes scaled by 1e12 and 1e-12, extreme
    parameter values: the NaN / inf paths of max(), min(), np.clip, np.interp, asof and the
    formatting agree with the reference."""
    from oracle.make_golden_classifier import run_reference
    ref = load_reference()
    for seed in range(500, 548):
        case = make_case(seed)
        rng = np.random.default_rng(seed)
        mode = seed % 5
        if mode == 0:
            idx = rng.choice(len(case["dev_values"]), size=max(1, len(case["dev_values"]) // 5), replace=False)
            case["dev_values"][idx] = np.nan
            case["dev_values"][0] = np.nan
        elif mode == 1:
            case["env"] = case["env"] * 1e12
        elif mode == 2:
            case["env"], case["floor"] = case["env"] * 1e-12, case["floor"] * 1e-12
        elif mode == 3:
            case["floor"][case["peaks"][::5]] = np.nan
        else:
            case["params"].update(min_bpm=10, max_bpm=400, stability_history_window=1, kickstart_check_threshold=1.1,
                                  cascade_reset_trigger_count=1, pairing_confidence_threshold=0.0)
        with np.errstate(all="ignore"):
            want = run_reference(ref, case)
        got, _ = run_ours(case, caplog)
        assert_same(got, want, f"seed {seed} mode {mode}")


def _job(case, with_text=True):
    hint = case["start_bpm"]
    packed = classifier.pack_params(case["params"], float(hint) if hint else 80.0, case["peak_time"],
                                    case["recovery_time"], with_text=with_text)
    return (case["env"], case["floor"], case["peaks"], case["dev_index"], case["dev_values"], case["rate"], packed)


def test_no_text_mode_and_threaded_batch_give_the_same_decisions():
    """BPM_CLASSIFY_NO_TEXT only drops the strings; bpm_classify_peaks_batch (one recording per host
    thread) returns for every job what the single call returns."""
    cases = [make_case(seed) for seed in range(40, 64)]
    single = [classifier.classify_arrays(*_job(c)) for c in cases]
    for n_threads in (1, 4, 0):
        batch = classifier.classify_batch([_job(c) for c in cases], n_threads=n_threads)
        for a, b in zip(single, batch):
            assert a["texts"] == b["texts"] and a["events"] == b["events"]
            for k in ("beat_positions", "peak_types", "history_times", "history_bpm"):
                assert np.array_equal(a[k], b[k]), k
    quiet = classifier.classify_batch([_job(c, with_text=False) for c in cases], n_threads=3)
    for a, b in zip(single, quiet):
        assert all(t == "" for t in b["texts"]) and len(b["texts"]) == len(a["texts"])
        assert a["events"] == b["events"] and a["final_long_term_bpm"] == b["final_long_term_bpm"]
        for k in ("beat_positions", "peak_types", "history_times", "history_bpm"):
            assert np.array_equal(a[k], b[k]), k
    bad = list(_job(cases[0]))
    bad[2] = bad[2][::-1].copy()                                        # descending peaks
    with pytest.raises(ValueError, match="job 1"):
        classifier.classify_batch([_job(cases[1]), tuple(bad)], n_threads=2)


def test_install_rebinds_classify_peaks():
    class Mod:
        class PeakClassifier:
            def classify_peaks(self):
                raise AssertionError("not replaced")
    classifier.install(Mod)
    assert Mod.PeakClassifier.classify_peaks is classifier.classify_peaks


def test_shipped_vulpine_run_labels_reproduced():
    """The reference's own shipped artefact for the sequential stage: samples/vulpine_Debug_Log.md
    labels each of the 1456 raw peaks of samples/vulpine (S1 / S2 / Lone S1 / Noise).  The compiled
    classifier + correction passes, chained as analyze_wav_file chains them (preliminary pass at
    0.75 -> start BPM and recovery window -> main pass -> stages 4 and 5) on the reference's own
    envelope / floor / raw peaks (tests/golden/vulpine.npz), must give every peak the shipped label
    and the 734 final beats."""
    import pandas as pd
    from bpm_analysis_b200 import corrections
    from bpm_analysis_b200.params import default_params
    from conftest import load_golden
    from oracle import ref_port
    g = load_golden("vulpine")
    with open(os.path.join(GOLDEN_DIR, "vulpine_labels.json")) as fh:
        shipped = json.load(fh)
    env, rate, peaks = g["envelope"], int(g["rate"]), g["raw_peaks"]
    assert np.allclose(np.array(shipped["times"]), peaks / rate, atol=6e-5) and len(shipped["labels"]) == len(peaks)
    params = default_params()
    base = {"env": env, "floor": g["floor"], "peaks": peaks, "dev_index": g["smoothed_dev_index"],
            "dev_values": g["smoothed_dev_values"], "rate": rate}
    pre = ClassifierStandIn(dict(base, params=dict(params, pairing_confidence_threshold=0.75), start_bpm=None,
                                 peak_time=None, recovery_time=None))
    anchors, _, _ = classifier.classify_peaks(pre)                      # _run_preliminary_pass, :1622-1653
    assert len(anchors) >= 10
    start_bpm = 60.0 / np.median(np.diff(anchors) / rate)
    series, times = ref_port.calculate_bpm_series(anchors, rate, params)
    peak_t = times[np.argmax(series.to_numpy())]                        # find_recovery_phase, :1612-1620
    main = ClassifierStandIn(dict(base, params=params, start_bpm=start_bpm, peak_time=peak_t,
                                  recovery_time=peak_t + 120.0))
    s1, raw, data = classifier.classify_peaks(main)
    floor = pd.Series(g["floor"], index=np.arange(len(env)))
    final = corrections.correct_peaks_by_rhythm(s1, env, rate, params)
    info = data["beat_debug_info"]
    for _ in range(5):
        final, info, made = corrections._fix_rhythmic_discontinuities(final, raw, info, env, floor, params, rate)
        if made == 0:
            break
    labels = [info[p].split("§")[0] for p in raw]
    assert labels == shipped["labels"]
    assert np.array_equal(final, g["beats"])
