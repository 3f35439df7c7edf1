"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (authoring container only).

Records what the UNMODIFIED reference writes to disk for one analysed recording
(``ReportGenerator`` bpm_analysis.py:782-983 and the CSV block of ``Plotter.plot_and_save``
:458-473) together with the inputs it was given, as ``tests/golden/reports_ramp.pkl.gz``:

    python oracle/make_report_golden.py

The CSV block lives inside a plotly method; its loop is executed here through the reference's own
``csv`` usage restated in :func:`reference_csv` (header row, ``f"{t:.3f}", f"{bpm:.3f}"`` per
non-NaN beat, csv.writer defaults).  The pickle holds pandas objects (pandas version recorded in
it); ``tests/test_reports_cpu.py`` replays the inputs through ``bpm_analysis_b200.reports``.
"""
from __future__ import annotations

import csv
import gzip
import io
import os
import pickle
import sys
import tempfile

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.load_reference import load_reference, reference_params   # noqa: E402


def reference_csv(final_metrics) -> bytes:
    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    w.writerow(["Time (s)", "Average BPM"])
    for t, bpm in zip(final_metrics["bpm_times"], final_metrics["smoothed_bpm"].values):
        if not np.isnan(bpm):
            w.writerow([f"{t:.3f}", f"{bpm:.3f}"])
    return buf.getvalue().encode("utf-8")


def analyse(ref, path, params, out_dir):
    """Stages 1-6 of analyze_wav_file (bpm_analysis.py:1731-1757) up to the metrics."""
    env, rate = ref.preprocess_audio(path, params, out_dir)
    floor, troughs = ref._calculate_dynamic_noise_floor(env, rate, params)
    sb, pt, rt = ref._run_preliminary_pass(env, rate, params, floor, troughs, None)
    clf = ref.PeakClassifier(env, rate, params, sb, floor, troughs, pt, rt)
    s1, raw, data = clf.classify_peaks()
    final, data = ref._refine_and_correct_peaks(s1, raw, data, env, rate, params)
    return env, rate, raw, data, ref._calculate_final_metrics(final, rate, params)


def reference_files(ref, path, out_dir, env, rate, raw, data, metrics, hint):
    rg = ref.ReportGenerator(path, out_dir)
    rg.save_analysis_summary(metrics)
    rg.create_chronological_log(env, rate, raw, data, metrics)
    rg.save_analysis_settings(hint)
    base = os.path.basename(os.path.splitext(path)[0])
    files = {s: open(os.path.join(out_dir, base + s), "rb").read()
             for s in ("_Analysis_Summary.md", "_Debug_Log.md", "_Analysis_Settings.json")}
    files["_bpm_plot.csv"] = reference_csv(metrics)
    return files


def main():
    from scipy.io import wavfile
    from bpm_analysis_b200 import synth
    ref = load_reference()
    params = reference_params()
    params["save_filtered_wav"] = False
    pcm, sr, _ = synth.pcg_recording(75.0, 8000, lambda t: 65.0 + 70.0 * np.exp(-((t - 35.0) / 14.0) ** 2), 77,
                                     noise_sigma=0.12)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ramp.wav")
        wavfile.write(path, sr, pcm)
        env, rate, raw, data, metrics = analyse(ref, path, params, d)
        files = reference_files(ref, path, d, env, rate, raw, data, metrics, 72.5)
    case = {"pandas": pd.__version__, "file_name": "ramp.wav", "envelope": env, "rate": rate, "raw_peaks": raw,
            "analysis_data": data, "final_metrics": metrics, "start_bpm_hint": 72.5, "files": files}
    out = os.path.join(ROOT, "tests", "golden", "reports_ramp.pkl.gz")
    with gzip.open(out, "wb") as fh:
        pickle.dump(case, fh, protocol=4)
    print(out, os.path.getsize(out), {k: len(v) for k, v in files.items()})


if __name__ == "__main__":
    main()
