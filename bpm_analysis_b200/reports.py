"""On-disk outputs of an analysis without plotly and without the
reference's row-by-row DataFrames.

What the reference leaves next to a recording and other tools read back:

=============================  ===============================  =================================
file                           reference writer                 here
=============================  ===============================  =================================
``<base>_bpm_plot.csv``        ``Plotter.plot_and_save``        :func:`write_bpm_plot_csv`
                               (bpm_analysis.py:458-473)
``<base>_Analysis_Summary.md`` ``ReportGenerator                :meth:`ReportGenerator.save_analysis_summary`
                               .save_analysis_summary``
                               (:802-814, :916-983)
``<base>_Debug_Log.md``        ``.create_chronological_log``    :meth:`ReportGenerator.create_chronological_log`
                               (:816-914)
``<base>_Analysis_Settings     ``.save_analysis_settings``      :meth:`ReportGenerator.save_analysis_settings`
.json``                        (:790-800)
=============================  ===============================  =================================

``heartbeat_labeler.py:30-116`` parses the debug log and the debug WAV, the hugging-face app
offers the summary and the CSV for download: the files must come out byte for byte as the
reference writes them (the two "generated on" time-stamp lines aside).  ``tests/test_reports_cpu.py``
compares them with the unmodified reference's files; ``install_reports`` rebinds
``bpm_analysis.ReportGenerator``; :func:`write_outputs` is the headless equivalent of stage 6 of
``analyze_wav_file`` (:1757-1765) for a batch service.

How it differs from the reference's implementation:

* the per-beat tables (CSV, heartbeat table) are formatted by ONE compiled call
  (``bpm_host_format_rows`` in libbpm_host.so) instead of one f-string per row;
* the debug log does not build the reference's M-row ``master_df`` and ``merge_asof`` it against
  the events (:840-855; 28.8 M rows x 3 columns for a 24-h recording): every event sits ON a
  sample of the envelope, so its row of the forward-filled frame is found by index -- the floor at
  the sample, and for the two sparse columns the last known value at or before the sample
  (``searchsorted``).  Same values, O(events log beats) instead of O(M).
"""
from __future__ import annotations

import ctypes as C
import datetime
import json
import logging
import os
import re
from typing import Dict, List, Optional

import numpy as np
import pandas as pd

from .classifier import load_host_library

SECTION = "§"                       # the reference joins the fields of a debug reason with it


def _fx(v, digits: int) -> str:
    return format(v, f".{digits}f")


# ------------------------------------------------------------------------------------------------
# tables
# ------------------------------------------------------------------------------------------------
def format_rows(a, b, prec_a: int, prec_b: int, head: str, mid: str, tail: str, skip_nan_b: bool = True) -> bytes:
    """``head + f"{a:.{prec_a}f}" + mid + f"{b:.{prec_b}f}" + tail`` for every row, rows with NaN ``b``
    left out, as UTF-8 bytes -- one compiled call."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = min(a.size, b.size)                                          # zip() stops at the shorter one
    if n == 0:
        return b""
    lib = load_host_library()
    h, m, t = head.encode(), mid.encode(), tail.encode()
    cap = n * (len(h) + len(m) + len(t) + 48)
    for _ in range(2):
        buf = C.create_string_buffer(cap)
        need = int(lib.bpm_host_format_rows(a.ctypes.data, b.ctypes.data, n, prec_a, prec_b, h, m, t,
                                            1 if skip_nan_b else 0, C.addressof(buf), cap))
        if need < 0:
            raise RuntimeError(f"bpm_host_format_rows failed ({need})")
        if need <= cap:
            return buf.raw[:need]
        cap = need
    raise RuntimeError("bpm_host_format_rows: size changed between calls")


def _series_rows(final_metrics: Dict):
    """(times, values) of the smoothed BPM series when the reference would print them, else None."""
    series, times = final_metrics.get("smoothed_bpm"), final_metrics.get("bpm_times")
    if series is None or series.empty or times is None:
        return None
    return np.asarray(times, dtype=np.float64), np.asarray(series.values, dtype=np.float64)


def bpm_plot_csv_bytes(final_metrics: Dict) -> Optional[bytes]:
    """The bytes of ``<base>_bpm_plot.csv`` (bpm_analysis.py:458-473), None when the reference writes no file."""
    rows = _series_rows(final_metrics)
    if rows is None:
        return None
    # csv.writer's default dialect: comma, "\r\n"; neither field ever needs quoting
    return b"Time (s),Average BPM\r\n" + format_rows(rows[0], rows[1], 3, 3, "", ",", "\r\n")


def write_bpm_plot_csv(file_name: str, output_directory: str, final_metrics: Dict) -> Optional[str]:
    data = bpm_plot_csv_bytes(final_metrics)
    if data is None:
        return None
    base = os.path.basename(os.path.splitext(file_name)[0])
    path = os.path.join(output_directory, f"{base}_bpm_plot.csv")
    try:
        with open(path, "wb") as fh:
            fh.write(data)
        logging.info(f"BPM plot data saved to {path}")
    except Exception as e:                                           # noqa: BLE001 - the reference logs and goes on
        logging.error(f"Failed to write BPM plot CSV: {e}")
        return None
    return path


# ------------------------------------------------------------------------------------------------
# debug reasons -> readable lines (Plotter.format_pairing_details_list / format_lone_s1_details_list,
# bpm_analysis.py:333-427; static, plotly-free, used by the debug log)
# ------------------------------------------------------------------------------------------------
_NUM_AT_END = re.compile(r"([\d\.]+)$")
_TIMES = re.compile(r"x([\d\.]+)")
_BY = re.compile(r"by ([\d\.]+)")
PAIRING_HEAD = "- S1-S2 pairing decision:"
LONE_HEAD = "- Lone S1 decision:"


def format_pairing_details_list(details_str: str) -> List[str]:
    rows = [ln.strip().lstrip("- ") for ln in details_str.strip().split("\n") if ln.strip()]
    if not rows:
        return [PAIRING_HEAD, "    - No details available."]
    out = [PAIRING_HEAD]
    try:
        m = _NUM_AT_END.search(rows[0])
        conf = float(m.group(1)) if m else 0.0
        out.append("    - " + rows[0])
        for row in rows[1:]:
            # each adjustment line carries its own factor / penalty; the running confidence is re-derived
            # from the text, not taken from the classifier
            if "Stability Pre-Adjust" in row:
                m = _TIMES.search(row)
                conf = conf * (float(m.group(1)) if m else 1)
                shown = conf
            elif "PENALIZED by" in row:
                m = _BY.search(row)
                conf = conf - (float(m.group(1)) if m else 0)
                shown = conf
            elif "Interval PENALTY by" in row:
                m = _BY.search(row)
                conf = conf - (float(m.group(1)) if m else 0)
                shown = max(0, conf)
            else:
                out.append("    - " + row)
                continue
            out.append(f"    - {row} -> {_fx(shown, 3)}")
    except (ValueError, IndexError):
        return [PAIRING_HEAD, f"    - {details_str}"]
    return out


_LONE = re.compile(r"(Validated|Rejected) Lone S1: Confidence ([\d\.]+) (>=|<) Threshold ([\d\.]+)\. \((.*)\)")
_LONE_PARTS = {
    "rhythm": re.compile(r"Rhythm Fit=([\d\.]+)"),
    "rhythm_note": re.compile(r"\(Interval .*?s vs Expected .*?s\)"),
    "amp": re.compile(r"Amplitude Fit=([\d\.]+)"),
    "amp_note": re.compile(r"\(Strength Ratio .*?x\)"),
    "w_rhythm": re.compile(r"Rhythm=([\d\.]+)"),
    "w_amp": re.compile(r"Amplitude=([\d\.]+)"),
}


def format_lone_s1_details_list(details_str: str) -> List[str]:
    unparsed = [LONE_HEAD, f"\t- {details_str}"]
    top = _LONE.search(details_str)
    if not top:
        return unparsed
    try:
        status, conf_s, op, thr_s, why = top.groups()
        conf, thr = float(conf_s), float(thr_s)
        hit = {k: rx.search(why) for k, rx in _LONE_PARTS.items()}
        rhythm = float(hit["rhythm"].group(1))
        out = [LONE_HEAD, f"\t- Rhythm Fit={_fx(rhythm, 2)} {hit['rhythm_note'].group(0)}"]
        amp = float(hit["amp"].group(1))
        out.append(f"\t- Amplitude Fit={_fx(amp, 2)} {hit['amp_note'].group(0)}")
        if hit["w_rhythm"] and hit["w_amp"]:
            wr, wa = float(hit["w_rhythm"].group(1)), float(hit["w_amp"].group(1))
            part_r, part_a = rhythm * wr, amp * wa
            out += ["\t- Weighted Calculation:",
                    f"\t\t- Rhythm: {_fx(rhythm, 2)} × {_fx(wr, 2)} = {_fx(part_r, 3)}",
                    f"\t\t- Amplitude: {_fx(amp, 2)} × {_fx(wa, 2)} = {_fx(part_a, 3)}",
                    f"\t\t- Final: {_fx(part_r, 3)} + {_fx(part_a, 3)} = {_fx(conf, 3)}"]
        verdict = "Validated" if "Validated" in status else "Rejected"
        out.append(f"- Final Score: Confidence {_fx(conf, 3)} {op} {_fx(thr, 2)} -> {verdict}")
    except (AttributeError, ValueError, IndexError) as e:
        logging.warning(f"Could not parse Lone S1 details string: '{details_str}'. Error: {e}")
        return unparsed
    return out


def _reason_lines(reason: str) -> List[str]:
    """The lines the log prints under a peak for its debug reason (bpm_analysis.py:868-895)."""
    if not reason or reason == "Unknown":
        return ["**Unclassified Peak**"]
    label, *fields = reason.split(SECTION)
    out = [f"**{label}.**"]
    for k in range(0, len(fields), 2):                               # (tag, value) pairs
        tag, value = fields[k], fields[k + 1] if k + 1 < len(fields) else ""
        if "PAIRING" in tag:
            out += format_pairing_details_list(value)
        elif "LONE_S1_REJECT_REASON" in tag or "LONE_S1_VALIDATE_REASON" in tag:
            out += format_lone_s1_details_list(value)
        elif "ORIGINAL_REASON" in tag:
            out.append(f"- Original Classification:\n    - `{value}`")
    return out


# ------------------------------------------------------------------------------------------------
# the forward-filled frame of the debug log, by index
# ------------------------------------------------------------------------------------------------
def _sparse_column(index_sec: np.ndarray, values: np.ndarray, rate, m: int):
    """A Series indexed by seconds, placed on the envelope's sample grid the way
    ``master_df[col] = series.groupby(level=0).mean()`` does (bpm_analysis.py:846-850): duplicate stamps
    averaged, stamps that are not EXACTLY ``k / rate`` for a sample k < m dropped by the index
    alignment.  -> (sorted sample positions, values) of the non-NaN entries (``ffill`` steps over NaN)."""
    if len(index_sec) == 0:
        return np.empty(0, np.int64), np.empty(0, np.float64)
    grouped = pd.Series(np.asarray(values, dtype=np.float64), index=np.asarray(index_sec, dtype=np.float64)
                        ).groupby(level=0).mean()
    stamps, vals = grouped.index.values, grouped.values
    with np.errstate(invalid="ignore"):
        k = np.rint(stamps * rate)
        ok = np.isfinite(k) & (k >= 0) & (k < m)
        k = np.where(ok, k, 0).astype(np.int64)
        ok &= (k / rate == stamps) & ~np.isnan(vals)
    return k[ok], vals[ok]


def _lookup_ffill(pos: np.ndarray, vals: np.ndarray, at: np.ndarray) -> np.ndarray:
    """Value of the forward-filled sparse column at samples ``at`` (NaN before its first entry)."""
    out = np.full(at.size, np.nan)
    if pos.size:
        j = np.searchsorted(pos, at, side="right") - 1
        out[j >= 0] = vals[j[j >= 0]]
    return out


def _dense_ffill_at(col: np.ndarray, at: np.ndarray) -> np.ndarray:
    col = np.asarray(col, dtype=np.float64)
    if not np.isnan(col).any():
        return col[at]
    last = np.where(~np.isnan(col), np.arange(col.size), -1)
    np.maximum.accumulate(last, out=last)
    j = last[at]
    out = np.full(at.size, np.nan)
    out[j >= 0] = col[j[j >= 0]]
    return out


def log_events(audio_envelope: np.ndarray, sample_rate, all_raw_peaks, analysis_data: Dict, smoothed_bpm, bpm_times
               ) -> Optional[Dict[str, object]]:
    """The merged event table of the debug log (``_prepare_log_data``, bpm_analysis.py:826-855) as plain
    arrays in chronological order: sample, time, is_peak, reason, amp and -- where the reference's frame
    has the column -- noise_floor / smoothed_bpm / lt_bpm (absent columns are ``None``)."""
    env = np.asarray(audio_envelope)
    m = env.size
    reasons_by_peak = analysis_data.get("beat_debug_info", {})
    samples, reasons, is_peak = [], [], []
    for p in all_raw_peaks:
        why = reasons_by_peak.get(p)
        if why:
            samples.append(int(p)), reasons.append(why), is_peak.append(True)
    if "trough_indices" in analysis_data:
        for p in analysis_data["trough_indices"]:
            samples.append(int(p)), reasons.append(""), is_peak.append(False)
    if not samples:
        return None
    at = np.asarray(samples, dtype=np.int64)
    times = at / sample_rate
    order = np.argsort(times, kind="stable")
    at, times = at[order], times[order]
    ev = {"sample": at, "time": times, "is_peak": np.asarray(is_peak)[order],
          "reason": [reasons[i] for i in order], "amp": env[at],
          "noise_floor": None, "smoothed_bpm": None, "lt_bpm": None}
    if "dynamic_noise_floor_series" in analysis_data:
        ev["noise_floor"] = _dense_ffill_at(analysis_data["dynamic_noise_floor_series"].values, at)
    if smoothed_bpm is not None and not smoothed_bpm.empty:
        ev["smoothed_bpm"] = _lookup_ffill(*_sparse_column(bpm_times, smoothed_bpm.values, sample_rate, m), at)
    lt = analysis_data.get("long_term_bpm_series")
    if lt is not None and not lt.empty:
        ev["lt_bpm"] = _lookup_ffill(*_sparse_column(lt.index.values, lt.values, sample_rate, m), at)
    return ev


_METRIC_ROWS = (("amp", "Raw Amp"), ("noise_floor", "Noise Floor"), ("smoothed_bpm", "Average BPM (Smoothed)"),
                ("lt_bpm", "Long-Term BPM (Belief)"))


def debug_log_text(file_name: str, ev: Optional[Dict[str, object]], now: Optional[datetime.datetime] = None) -> str:
    """The text of ``<base>_Debug_Log.md`` (``_write_log_events``, bpm_analysis.py:857-914)."""
    if ev is None or len(ev["sample"]) == 0:
        return "# No significant events detected to log.\n"
    now = now or datetime.datetime.now()
    parts = [f"# Chronological Debug Log for {os.path.basename(file_name)}\n",
             f"Analysis performed on: {now.strftime('%Y-%m-%d %H:%M:%S')}\n\n"]
    cols = [(title, ev[key]) for key, title in _METRIC_ROWS if ev[key] is not None]
    shown = [(title, [None if np.isnan(v) else _fx(v, 1) for v in col]) for title, col in cols]
    stamps = [_fx(t, 4) for t in ev["time"]]
    for i, stamp in enumerate(stamps):
        parts.append(f"## Time: `{stamp}s`\n")
        if ev["is_peak"][i]:
            parts.append("\n".join(_reason_lines(ev["reason"][i])) + "\n")
        else:
            parts.append("**Trough Detected**\n")
        for title, col in shown:
            if col[i] is not None:
                parts.append(f"- **{title}**: `{col[i]}`\n")
        parts.append("\n\n")
    return "".join(parts)


# ------------------------------------------------------------------------------------------------
# summary
# ------------------------------------------------------------------------------------------------
def _seconds_since_epoch(stamp) -> float:
    return (stamp - datetime.datetime.fromtimestamp(0)).total_seconds()


def _slope_block(stats: Optional[Dict], sign: str, none_text: str) -> str:
    if not stats:
        return none_text + "\n\n"
    span = f"{stats['start_time'].strftime('%M:%S')} to {stats['end_time'].strftime('%M:%S')}"
    return ("| Attribute | Value |\n|:---|:---|\n"
            f"| **Rate** | `{sign}{_fx(stats['slope_bpm_per_sec'], 2)}` BPM/second |\n"
            f"| **Period** | {span} |\n"
            f"| **Duration** | {_fx(stats['duration_sec'], 1)} seconds |\n"
            f"| **BPM Change** | {_fx(stats['start_bpm'], 1)} to {_fx(stats['end_bpm'], 1)} BPM |\n\n")


def _change_list(changes, key: str, sign: str) -> str:
    if not changes:
        return "*None found.*\n"
    return "".join(f"- **From {_fx(_seconds_since_epoch(c['start_time']), 1)}s to "
                   f"{_fx(_seconds_since_epoch(c['end_time']), 1)}s:** Duration={_fx(c['duration_sec'], 1)}s, "
                   f"Change=`{sign}{_fx(c[key], 1)}` BPM\n" for c in changes)


def summary_bytes(file_name: str, final_metrics: Dict, now: Optional[datetime.datetime] = None) -> bytes:
    """The bytes of ``<base>_Analysis_Summary.md`` (bpm_analysis.py:802-814 and :916-983)."""
    now = now or datetime.datetime.now()
    hrv, hrr = final_metrics.get("hrv_summary"), final_metrics.get("hrr_stats")
    s = [f"# Analysis Report for: {os.path.basename(file_name)}\n",
         f"*Generated on: {now.strftime('%Y-%m-%d %H:%M:%S')}*\n\n",
         "## Overall Summary\n\n| Metric | Value |\n|:---|:---|\n"]
    if hrv:
        if hrv.get("avg_bpm") is not None:
            s.append(f"| **Average BPM** | {_fx(hrv['avg_bpm'], 1)} BPM |\n")
            s.append(f"| **BPM Range** | {_fx(hrv['min_bpm'], 1)} to {_fx(hrv['max_bpm'], 1)} BPM |\n")
        if hrv.get("avg_rmssdc") is not None:
            s.append(f"| **Avg. Corrected RMSSD** | {_fx(hrv['avg_rmssdc'], 2)} |\n")
        if hrv.get("avg_sdnn") is not None:
            s.append(f"| **Avg. Windowed SDNN** | {_fx(hrv['avg_sdnn'], 2)} ms |\n")
    if hrr and hrr.get("hrr_value_bpm") is not None:
        s.append(f"| **1-Minute HRR** | {_fx(hrr['hrr_value_bpm'], 1)} BPM Drop |\n")
    s.append("\n## Steepest Slopes Analysis\n\n### Peak Exertion (Fastest HR Increase)\n\n")
    s.append(_slope_block(final_metrics.get("peak_exertion_stats"), "+", "*No significant peak exertion period found.*"))
    s.append("### Peak Recovery (Fastest HR Decrease)\n\n")
    s.append(_slope_block(final_metrics.get("peak_recovery_stats"), "",
                          "*No significant peak recovery period found post-peak.*"))
    s.append("## All Significant HR Changes\n\n### Exertion Periods (Sustained HR Increase)\n\n")
    s.append(_change_list(final_metrics.get("major_inclines"), "bpm_increase", "+"))
    s.append("\n### Recovery Periods (Sustained HR Decrease)\n\n")
    s.append(_change_list(final_metrics.get("major_declines"), "bpm_decrease", "-"))
    s.append("\n## Heartbeat Data (BPM over Time)\n\n| Time (s) | Average BPM |\n|:---:|:---:|\n")
    text = "".join(s).encode("utf-8")
    rows = _series_rows(final_metrics)
    if rows is None:
        return text + b"| *No data* | *No data* |\n"
    return text + format_rows(rows[0], rows[1], 2, 1, "| ", " | ", " |\n")


# ------------------------------------------------------------------------------------------------
# the reference's class, same constructor and method names
# ------------------------------------------------------------------------------------------------
class ReportGenerator:
    """Drop-in for ``bpm_analysis.ReportGenerator`` (bpm_analysis.py:782-983)."""

    def __init__(self, file_name: str, output_directory: str):
        self.file_name = file_name
        self.output_directory = output_directory
        self.file_name_no_ext = os.path.splitext(file_name)[0]
        self.base_name = os.path.basename(self.file_name_no_ext)

    def _path(self, suffix: str) -> str:
        return os.path.join(self.output_directory, self.base_name + suffix)

    def save_analysis_settings(self, start_bpm_hint: Optional[float]):
        path = self._path("_Analysis_Settings.json")
        try:
            with open(path, "w", encoding="utf-8") as fh:
                json.dump({"start_bpm_hint": start_bpm_hint}, fh, indent=4)
            logging.info(f"Analysis settings saved to {path}")
        except Exception as e:                                       # noqa: BLE001 - as the reference: log, go on
            logging.error(f"Could not save analysis settings file. Error: {e}")

    def save_analysis_summary(self, final_metrics: Dict):
        path = self._path("_Analysis_Summary.md")
        with open(path, "wb") as fh:
            fh.write(summary_bytes(self.file_name, final_metrics))
        logging.info(f"Markdown analysis summary saved to {path}")

    def create_chronological_log(self, audio_envelope: np.ndarray, sample_rate: int, all_raw_peaks: np.ndarray,
                                 analysis_data: Dict, final_metrics: Dict):
        path = self._path("_Debug_Log.md")
        logging.info(f"Generating readable debug log at '{path}'...")
        ev = log_events(audio_envelope, sample_rate, all_raw_peaks, analysis_data, final_metrics.get("smoothed_bpm"),
                        final_metrics.get("bpm_times"))
        with open(path, "w", encoding="utf-8") as fh:
            fh.write(debug_log_text(self.file_name, ev))
        logging.info("Debug log generation complete.")


def write_outputs(file_name: str, output_directory: str, audio_envelope: np.ndarray, sample_rate: int,
                  all_raw_peaks: np.ndarray, analysis_data: Dict, final_metrics: Dict,
                  start_bpm_hint: Optional[float] = None) -> Dict[str, Optional[str]]:
    """Stage 6 of ``analyze_wav_file`` (bpm_analysis.py:1757-1765) without the plotly figure: CSV, summary,
    debug log and settings of one analysed recording.  -> the paths written."""
    rep = ReportGenerator(file_name, output_directory)
    csv_path = write_bpm_plot_csv(file_name, output_directory, final_metrics)
    rep.save_analysis_summary(final_metrics)
    rep.create_chronological_log(audio_envelope, sample_rate, all_raw_peaks, analysis_data, final_metrics)
    rep.save_analysis_settings(start_bpm_hint)
    return {"csv": csv_path, "summary": rep._path("_Analysis_Summary.md"), "debug_log": rep._path("_Debug_Log.md"),
            "settings": rep._path("_Analysis_Settings.json")}


def install_reports(ref_module):
    """Rebind ``ReportGenerator`` on an imported reference module (its ``Plotter`` keeps writing the
    HTML figure and, in passing, the CSV -- plotly territory)."""
    ref_module.ReportGenerator = ReportGenerator
    return ref_module
