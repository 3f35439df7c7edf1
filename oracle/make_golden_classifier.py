"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/classifier_cases.json.gz``.

Run in the authoring container (needs ``/root/reference``):

    python -m oracle.make_golden_classifier

For each seeded case of ``tests/_classifier_cases.py`` the UNMODIFIED reference
``PeakClassifier.classify_peaks`` (bpm_analysis.py:113-131) is run on a classifier object whose
state was filled in directly (the constructor's numeric work is not what is being recorded), and
its candidate beats, per-peak debug strings, long-term BPM trace (hex floats) and log lines are
stored.  ``tests/test_classifier_cpu.py`` requires libbpm_host.so to reproduce them exactly.
"""
from __future__ import annotations

import gzip
import json
import logging
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from _classifier_cases import make_case, pack_result, state_of   # noqa: E402
from oracle.load_reference import load_reference               # noqa: E402

GOLDEN_SEEDS = list(range(24))
OUT = os.path.join(REPO, "tests", "golden", "classifier_cases.json.gz")


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__(level=logging.INFO)
        self.lines = []

    def emit(self, record):
        self.lines.append(record.getMessage())


def reference_classifier(ref, case):
    obj = object.__new__(ref.PeakClassifier)
    obj.audio_envelope, obj.sample_rate, obj.params = case["env"], case["rate"], case["params"]
    obj.peak_bpm_time_sec, obj.recovery_end_time_sec = case["peak_time"], case["recovery_time"]
    obj.state = state_of(case)
    return obj


def run_reference(ref, case):
    cap = _Capture()
    root = logging.getLogger()
    level = root.level
    root.addHandler(cap)
    root.setLevel(logging.INFO)
    try:
        obj = reference_classifier(ref, case)
        packed = pack_result(obj.classify_peaks())
    finally:
        root.removeHandler(cap)
        root.setLevel(level)
    packed["log"] = cap.lines
    packed["final_long_term_bpm"] = float(obj.state["long_term_bpm"]).hex()
    packed["final_consecutive_rr_rejections"] = int(obj.state["consecutive_rr_rejections"])
    return packed


def run_reference_corrections(ref, case, classified):
    """The reference's stages 4 and 5 (``_refine_and_correct_peaks``, bpm_analysis.py:1655-1698) on
    the classifier's result, with every intermediate recorded."""
    import numpy as np
    import pandas as pd
    cap = _Capture()
    root = logging.getLogger()
    level = root.level
    root.addHandler(cap)
    root.setLevel(logging.INFO)
    try:
        raw = case["peaks"]
        keys = list(raw)
        info = {keys[raw.tolist().index(k)]: t for k, t in zip(classified["keys"], classified["texts"])}
        s1 = np.array(classified["final_peaks"], dtype=np.int64)
        floor = pd.Series(case["floor"], index=np.arange(len(case["floor"])))
        peaks = ref.correct_peaks_by_rhythm(s1, case["env"], case["rate"], case["params"])
        out = {"after_rhythm": [int(x) for x in peaks], "iterations": []}
        for _ in range(5):
            peaks, info, made = ref._fix_rhythmic_discontinuities(peaks, raw, info, case["env"], floor, case["params"],
                                                                  case["rate"])
            out["iterations"].append({"peaks": [int(x) for x in peaks], "made": int(made)})
            if made == 0:
                break
        out["final_keys"] = [int(k) for k in info.keys()]
        out["final_texts"] = list(info.values())
    finally:
        root.removeHandler(cap)
        root.setLevel(level)
    out["log"] = cap.lines
    return out


def shipped_vulpine_labels():
    """The label the reference's OWN shipped run gave every raw peak of ``samples/vulpine``:
    ``samples/vulpine_Debug_Log.md`` lists each raw peak (and trough) in time order with its label
    in bold.  -> tests/golden/vulpine_labels.json"""
    import re
    from oracle.load_reference import REFERENCE_ROOT
    log = open(os.path.join(REFERENCE_ROOT, "samples", "vulpine_Debug_Log.md"), encoding="utf-8").read()
    entries = re.findall(r"## Time: `([0-9.]+)s`\n\*\*(.+?)\*\*", log)
    peaks = [(float(t), label.rstrip(".")) for t, label in entries if "Trough" not in label]
    out = os.path.join(REPO, "tests", "golden", "vulpine_labels.json")
    with open(out, "w") as fh:
        json.dump({"source": "samples/vulpine_Debug_Log.md", "times": [t for t, _ in peaks],
                   "labels": [l for _, l in peaks]}, fh)
    print(f"{out}: {len(peaks)} raw-peak labels")


def main():
    shipped_vulpine_labels()
    ref = load_reference()
    out = {}
    for seed in GOLDEN_SEEDS:
        case = make_case(seed)
        out[str(seed)] = run_reference(ref, case)
        out[str(seed)]["corrections"] = run_reference_corrections(ref, case, out[str(seed)])
    with gzip.GzipFile(OUT, "wb", mtime=0) as fh:
        fh.write(json.dumps(out, ensure_ascii=False, sort_keys=True).encode("utf-8"))
    n = sum(len(v["keys"]) for v in out.values())
    print(f"{OUT}: {len(out)} cases, {n} classified peaks, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
