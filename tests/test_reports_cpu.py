"""On-disk outputs: `_bpm_plot.csv`, `_Analysis_Summary.md`,
`_Debug_Log.md`, `_Analysis_Settings.json` byte for byte as the UNMODIFIED reference writes them
(bpm_analysis.py:458-473, :782-983), the two time-stamp lines aside.

* golden: inputs + files recorded from the reference by ``oracle/make_report_golden.py``;
* live (authoring container only): the reference analyses recordings here and both writers run on
  the same objects; the plotly-free text helpers are compared on adversarial strings."""
import gzip
import os
import pickle
import re

import numpy as np
import pandas as pd
import pytest

from bpm_analysis_b200 import reports
from oracle.load_reference import load_reference, reference_available

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reports_ramp.pkl.gz")
STAMP = re.compile(rb"(Generated on: |Analysis performed on: )[0-9: -]+")
SUFFIXES = ("_Analysis_Summary.md", "_Debug_Log.md", "_Analysis_Settings.json", "_bpm_plot.csv")


def unstamped(b: bytes) -> bytes:
    return STAMP.sub(rb"\1<now>", b)


def written(out_dir, base):
    return {s: open(os.path.join(out_dir, base + s), "rb").read() for s in SUFFIXES}


def test_golden_files_byte_for_byte(tmp_path):
    with gzip.open(GOLDEN, "rb") as fh:
        case = pickle.load(fh)
    if case["pandas"] != pd.__version__:
        pytest.skip(f"fixture pickled with pandas {case['pandas']}")
    paths = reports.write_outputs(str(tmp_path / case["file_name"]), str(tmp_path), case["envelope"], case["rate"],
                                  case["raw_peaks"], case["analysis_data"], case["final_metrics"],
                                  case["start_bpm_hint"])
    assert all(os.path.isfile(p) for p in paths.values())
    got = written(str(tmp_path), "ramp")
    assert len(case["files"]["_Debug_Log.md"]) > 100_000
    for s in SUFFIXES:
        assert unstamped(got[s]) == unstamped(case["files"][s]), s
    # what heartbeat_labeler.py:36-41 reads back: pd.read_csv of the BPM table, by column name
    table = pd.read_csv(paths["csv"])
    fm = case["final_metrics"]
    assert list(table.columns) == ["Time (s)", "Average BPM"] and len(table) == int((~np.isnan(fm["smoothed_bpm"].values)).sum())
    assert np.allclose(table["Time (s)"].values, fm["bpm_times"], atol=5.1e-4, rtol=0)
    assert np.allclose(table["Average BPM"].values, fm["smoothed_bpm"].values, atol=5.1e-4, rtol=0)
    # one "## Time" header per logged event
    n_events = got["_Debug_Log.md"].count(b"## Time: `")
    n_peaks = sum(1 for p in case["raw_peaks"] if case["analysis_data"]["beat_debug_info"].get(p))
    assert n_events == n_peaks + len(case["analysis_data"]["trough_indices"])


def test_format_rows_matches_python_formatting():
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.uniform(-1e4, 1e4, 4000), rng.uniform(0, 1, 500) * 10.0 ** rng.integers(-8, 14, 500),
                        [0.0005, 0.0015, 0.0025, 2.5, 0.125, 1e300, -0.0, np.inf, -np.inf, np.nan, 0.05, 0.15, 0.25,
                         1e15 + 0.5, 123456.7895, 2.675, 1.005]])
    b = np.concatenate([rng.uniform(-300, 300, a.size - 40), np.full(20, np.nan), np.round(rng.uniform(0, 9, 20), 1) + 0.05])
    for pa, pb, head, mid, tail in ((3, 3, "", ",", "\r\n"), (2, 1, "| ", " | ", " |\n"), (4, 0, "<", ";", ">")):
        want = "".join(f"{head}{x:.{pa}f}{mid}{y:.{pb}f}{tail}" for x, y in zip(a, b) if not np.isnan(y))
        assert reports.format_rows(a, b, pa, pb, head, mid, tail) == want.encode()
    keep = "".join(f"{x:.1f},{y:.1f}\n" for x, y in zip(a, b))
    assert reports.format_rows(a, b, 1, 1, "", ",", "\n", skip_nan_b=False) == keep.encode()
    assert reports.format_rows(a[:0], b[:0], 1, 1, "", ",", "\n") == b""
    assert reports.format_rows(a, b[:7], 1, 1, "", ",", "\n") == "".join(f"{x:.1f},{y:.1f}\n" for x, y in zip(a, b[:7])).encode()


def test_format_rows_argument_errors_and_resize():
    """The C entry never throws: bad arguments come back as BPM_HOST_ERR_ARG (-1), a buffer that is too
    small as the size to retry with (include/bpm_host.h)."""
    import ctypes as C
    from bpm_analysis_b200.classifier import load_host_library
    lib = load_host_library()
    a = np.array([1.0, 2.5, 1e300]); b = np.array([3.0, np.nan, 4.0])
    pa, pb = a.ctypes.data, b.ctypes.data
    buf = C.create_string_buffer(8)
    call = lib.bpm_host_format_rows
    assert call(None, pb, 3, 1, 1, b"", b",", b"\n", 1, C.addressof(buf), 8) == -1
    assert call(pa, pb, -1, 1, 1, b"", b",", b"\n", 1, C.addressof(buf), 8) == -1
    assert call(pa, pb, 3, 18, 1, b"", b",", b"\n", 1, C.addressof(buf), 8) == -1
    assert call(pa, pb, 3, 1, 1, None, b",", b"\n", 1, C.addressof(buf), 8) == -1
    assert call(pa, pb, 3, 1, 1, b"", b",", b"\n", 1, None, 8) == -1
    want = "".join(f"{x:.1f},{y:.1f}\n" for x, y in zip(a, b) if not np.isnan(y)).encode()
    need = call(pa, pb, 3, 1, 1, b"", b",", b"\n", 1, None, 0)            # size query
    assert need == len(want) > 300                                        # 1e300 prints 301 digits
    assert call(pa, pb, 3, 1, 1, b"", b",", b"\n", 1, C.addressof(buf), 8) == need
    big = C.create_string_buffer(need)
    assert call(pa, pb, 3, 1, 1, b"", b",", b"\n", 1, C.addressof(big), need) == need and big.raw == want
    assert reports.format_rows(a, b, 1, 1, "", ",", "\n") == want          # the wrapper retries with the size
    assert call(pa, pb, 0, 1, 1, b"", b",", b"\n", 1, None, 0) == 0


def test_empty_and_degenerate_outputs(tmp_path):
    empty = {"smoothed_bpm": pd.Series(dtype=float), "bpm_times": np.array([]), "hrv_summary": {}, "hrr_stats": None,
             "major_inclines": [], "major_declines": [], "peak_recovery_stats": None, "peak_exertion_stats": None,
             "windowed_hrv_df": pd.DataFrame()}
    assert reports.bpm_plot_csv_bytes(empty) is None
    assert reports.write_bpm_plot_csv("x.wav", str(tmp_path), empty) is None
    text = reports.summary_bytes("x.wav", empty).decode()
    assert text.endswith("| *No data* | *No data* |\n") and "*None found.*" in text
    assert "*No significant peak exertion period found.*" in text
    assert reports.log_events(np.zeros(10), 10, np.array([], dtype=np.int64), {}, None, None) is None
    assert reports.debug_log_text("x.wav", None) == "# No significant events detected to log.\n"
    # a NaN stretch of the dense column is forward-filled, a sparse column is NaN before its first stamp
    col = np.array([np.nan, 1.0, np.nan, np.nan, 4.0, np.nan])
    assert np.array_equal(reports._dense_ffill_at(col, np.arange(6)), [np.nan, 1, 1, 1, 4, 4], equal_nan=True)
    pos, val = reports._sparse_column(np.array([0.2, 0.2, 0.5, 0.55]), np.array([1.0, 3.0, 7.0, 9.0]), 10, 6)
    assert pos.tolist() == [2, 5] and val.tolist() == [2.0, 7.0]          # 0.55 is no sample stamp, 0.2 averaged
    assert np.array_equal(reports._lookup_ffill(pos, val, np.arange(6)), [np.nan, np.nan, 2, 2, 2, 7], equal_nan=True)


needs_reference = pytest.mark.skipif(not reference_available(), reason="needs /root/reference")


@needs_reference
def test_text_helpers_match_the_reference_on_adversarial_strings():
    ref = load_reference()
    pairing = ["", "   \n  ", "Base Pairing Confidence: 0.62", "- Base 0.5\n- Stability Pre-Adjust x1.20\n- PENALIZED by 0.30",
               "conf 0.4\nInterval PENALTY by 0.9\nPENALIZED by 0.1\nnote", "no number here\nPENALIZED by ...",
               "x 1.2.3", "- - -  lead 0.75\n\n- Stability Pre-Adjust (none)\nInterval PENALTY by 0.10 then by 0.2",
               "a 1.\nStability Pre-Adjust x.\n", "only 7\nStability Pre-Adjust x2 PENALIZED by 1"]
    for s in pairing:
        assert reports.format_pairing_details_list(s) == ref.Plotter.format_pairing_details_list(s), repr(s)
    full = ("Validated Lone S1: Confidence 0.712 >= Threshold 0.55. (Rhythm Fit=0.81 (Interval 0.612s vs Expected 0.640s), "
            "Amplitude Fit=0.55 (Strength Ratio 1.32x), Weights: Rhythm=0.65, Amplitude=0.35)")
    lone = ["", "garbage", full, full.replace("Validated", "Rejected").replace(">=", "<"),
            full.replace(", Weights: Rhythm=0.65, Amplitude=0.35", ""), full.replace("Rhythm Fit=0.81", "Rhythm Fit=."),
            full.replace("(Strength Ratio 1.32x)", ""), full.replace("Amplitude Fit=0.55", "Amp=0.55"),
            full.replace("Rhythm=0.65", "Rhythm=1.2.3"), "Rejected Lone S1: Confidence 0.1 < Threshold 0.5. ()"]
    for s in lone:
        assert reports.format_lone_s1_details_list(s) == ref.Plotter.format_lone_s1_details_list(s), repr(s)


@needs_reference
@pytest.mark.parametrize("case", ["ramp", "c1_short", "vulpine", "too_quiet"])
def test_live_files_match_the_reference(case, tmp_path):
    from scipy.io import wavfile
    from bpm_analysis_b200 import synth
    from oracle.load_reference import REFERENCE_ROOT, reference_params
    from oracle.make_report_golden import analyse, reference_files
    ref = load_reference()
    params = reference_params()
    params["save_filtered_wav"] = False
    if case == "ramp":
        pcm, sr, _ = synth.pcg_recording(150.0, 8000, lambda t: 65.0 + 70.0 * np.exp(-((t - 60.0) / 25.0) ** 2), 77,
                                         noise_sigma=0.12)
    elif case == "c1_short":
        pcm, sr, _ = synth.config_c1(seed=5, duration_sec=45.0)
    elif case == "vulpine":
        sr, pcm = wavfile.read(os.path.join(REFERENCE_ROOT, "samples", "vulpine_filtered_debug.wav"))
    else:
        pcm, sr, _ = synth.pcg_recording(30.0, 4000, lambda t: 70.0 + 0 * t, 9, noise_sigma=0.9)
    (tmp_path / "ref").mkdir(), (tmp_path / "mine").mkdir()
    path = str(tmp_path / "rec.wav")
    wavfile.write(path, sr, pcm)
    env, rate, raw, data, metrics = analyse(ref, path, params, str(tmp_path / "ref"))
    hint = None if case == "vulpine" else 80
    want = reference_files(ref, path, str(tmp_path / "ref"), env, rate, raw, data, metrics, hint)
    reports.write_outputs(path, str(tmp_path / "mine"), env, rate, raw, data, metrics, hint)
    got = written(str(tmp_path / "mine"), "rec")
    for s in SUFFIXES:
        assert unstamped(got[s]) == unstamped(want[s]), s
    # install_reports: the reference's own stage-6 calls (bpm_analysis.py:1763-1766) through the rebound class
    import importlib
    try:
        reports.install_reports(ref)
        (tmp_path / "inst").mkdir()
        rg = ref.ReportGenerator(path, str(tmp_path / "inst"))
        assert isinstance(rg, reports.ReportGenerator)
        rg.save_analysis_summary(metrics)
        rg.create_chronological_log(env, rate, raw, data, metrics)
        rg.save_analysis_settings(hint)
        for s in SUFFIXES[:3]:
            assert unstamped(open(str(tmp_path / "inst" / ("rec" + s)), "rb").read()) == unstamped(want[s]), s
    finally:
        importlib.reload(ref)
