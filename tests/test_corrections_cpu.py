"""The compiled correction passes (libbpm_host.so) against the reference's
``correct_peaks_by_rhythm`` / ``_fix_rhythmic_discontinuities`` (bpm_analysis.py:1257-1412), run as
``_refine_and_correct_peaks`` runs them (:1655-1698): identical peaks after every pass, identical
debug-string dict, identical correction counts and log lines.  Golden outputs of the unmodified
reference: tests/golden/classifier_cases.json.gz (oracle/make_golden_classifier.py); live
comparison on more seeds where the reference tree is present.
"""
import gzip
import json
import logging
import os

import numpy as np
import pandas as pd
import pytest

from _classifier_cases import ClassifierStandIn, make_case
from bpm_analysis_b200 import classifier, corrections
from conftest import GOLDEN_DIR
from oracle.load_reference import load_reference, reference_available

needs_reference = pytest.mark.skipif(not reference_available(), reason="needs /root/reference")


def run_ours(case, caplog):
    final_peaks, raw, data = classifier.classify_peaks(ClassifierStandIn(case))
    info = data["beat_debug_info"]
    floor = pd.Series(case["floor"], index=np.arange(len(case["floor"])))
    caplog.clear()
    with caplog.at_level(logging.INFO):
        peaks = corrections.correct_peaks_by_rhythm(final_peaks, case["env"], case["rate"], case["params"])
        out = {"after_rhythm": [int(x) for x in peaks], "iterations": []}
        for _ in range(5):
            peaks, info, made = corrections._fix_rhythmic_discontinuities(peaks, raw, info, case["env"], floor,
                                                                          case["params"], case["rate"])
            assert peaks.dtype == np.int64
            out["iterations"].append({"peaks": [int(x) for x in peaks], "made": int(made)})
            if made == 0:
                break
    out["final_keys"] = [int(k) for k in info.keys()]
    out["final_texts"] = list(info.values())
    out["log"] = [r.getMessage() for r in caplog.records]
    return out


def assert_same(got, want, label):
    assert got["after_rhythm"] == want["after_rhythm"], label
    assert got["iterations"] == want["iterations"], label
    assert got["final_keys"] == want["final_keys"], label
    assert got["final_texts"] == want["final_texts"], label
    assert got["log"] == want["log"], label


def test_golden_corrections_reproduced_exactly(caplog):
    with gzip.open(os.path.join(GOLDEN_DIR, "classifier_cases.json.gz"), "rb") as fh:
        golden = json.loads(fh.read().decode("utf-8"))
    relabels = removals = conflicts = 0
    for seed, want in sorted(golden.items(), key=lambda kv: int(kv[0])):
        got = run_ours(make_case(int(seed)), caplog)
        assert_same(got, want["corrections"], f"seed {seed}")
        log = want["corrections"]["log"]
        relabels += sum("Re-labeling" in l for l in log)
        removals += sum("Removing weaker" in l for l in log)
        conflicts += sum("Conflict at" in l for l in log)
    assert relabels >= 20 and removals >= 10 and conflicts >= 20      # every branch is exercised


@needs_reference
def test_live_reference_corrections_many_seeds(caplog):
    from oracle.make_golden_classifier import run_reference, run_reference_corrections
    ref = load_reference()
    for seed in range(300, 380):
        case = make_case(seed)
        want = run_reference_corrections(ref, case, run_reference(ref, case))
        assert_same(run_ours(case, caplog), want, f"seed {seed}")


def test_short_inputs_are_returned_untouched(caplog):
    case = make_case(1)
    few = case["peaks"][:4]
    assert corrections.correct_peaks_by_rhythm(few, case["env"], 300, case["params"]) is few      # :1263-1264
    info = {}
    floor = pd.Series(case["floor"])
    with caplog.at_level(logging.INFO):
        peaks, out_info, made = corrections._fix_rhythmic_discontinuities(case["peaks"][:5], case["peaks"], info,
                                                                          case["env"], floor, case["params"], 300)
    assert peaks is not None and out_info is info and made == 0                                   # :1319-1321
    assert "Skipping correction pass" in caplog.records[-1].getMessage()


def test_install_rebinds_correction_passes():
    class Mod:
        pass
    corrections.install(Mod)
    assert Mod.correct_peaks_by_rhythm is corrections.correct_peaks_by_rhythm
    assert Mod._fix_rhythmic_discontinuities is corrections._fix_rhythmic_discontinuities
