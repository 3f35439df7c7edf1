"""install(): the reference module's names are rebound to the GPU mirror with identical
signatures.  Runs only where the reference tree is present (the authoring container)."""
import inspect

import pytest

from oracle.load_reference import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs /root/reference")


def test_install_rebinds_with_identical_signatures():
    pytest.importorskip("torch")
    import importlib
    ref = load_reference()
    from bpm_analysis_b200 import frontend
    names = ["preprocess_audio", "_calculate_dynamic_noise_floor", "calculate_bpm_series", "find_peak_recovery_rate",
             "find_peak_exertion_rate", "find_major_hr_inclines", "find_major_hr_declines", "calculate_windowed_hrv"]
    before = {n: inspect.signature(getattr(ref, n)) for n in names}
    m_before = {n: inspect.signature(getattr(ref.PeakClassifier, n)) for n in ("_find_raw_peaks", "_initialize_state")}
    try:
        frontend.install(ref)
        for n in names:
            assert getattr(ref, n) is getattr(frontend, n)
            a, b = before[n], inspect.signature(getattr(ref, n))
            assert list(a.parameters) == list(b.parameters), n
            assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()], n
        for n, sig in m_before.items():
            assert list(sig.parameters) == list(inspect.signature(getattr(ref.PeakClassifier, n)).parameters)
    finally:
        importlib.reload(ref)


def test_hot_path_params_match_reference_defaults():
    from oracle.load_reference import reference_params
    from bpm_analysis_b200.params import HOT_PATH_DEFAULTS
    rp = reference_params()
    for k, v in HOT_PATH_DEFAULTS.items():
        assert rp[k] == v, k


def test_sequential_stage_drop_ins_through_the_reference_orchestrator(tmp_path):
    """classifier.install + corrections.install on the reference module, then the reference's OWN
    stages 2-5 (`_run_preliminary_pass`, main `PeakClassifier`, `_refine_and_correct_peaks`,
    bpm_analysis.py:1622-1760) on a recording: same start BPM / recovery window, same final peaks,
    same debug dict as the untouched reference (the front end stays the reference's CPU code here)."""
    import importlib
    import numpy as np
    from scipy.io import wavfile
    from bpm_analysis_b200 import classifier, corrections, synth
    from oracle.load_reference import reference_params
    ref = load_reference()
    params = reference_params()
    params["save_filtered_wav"] = False
    pcm, sr, _ = synth.pcg_recording(150.0, 8000, lambda t: 65.0 + 70.0 * np.exp(-((t - 60.0) / 25.0) ** 2), 77,
                                     noise_sigma=0.12)
    path = str(tmp_path / "x.wav")
    wavfile.write(path, sr, pcm)
    env, rate = ref.preprocess_audio(path, params, str(tmp_path))
    floor, troughs = ref._calculate_dynamic_noise_floor(env, rate, params)

    def stages_2_to_5():
        start_bpm, peak_t, rec_t = ref._run_preliminary_pass(env, rate, params, floor, troughs, None)
        clf = ref.PeakClassifier(env, rate, params, start_bpm, floor, troughs, peak_t, rec_t)
        s1, raw, data = clf.classify_peaks()
        final, data = ref._refine_and_correct_peaks(s1, raw, data, env, rate, params)
        return start_bpm, peak_t, rec_t, s1, raw, final, data

    want = stages_2_to_5()
    sigs = {n: inspect.signature(getattr(ref, n)) for n in ("correct_peaks_by_rhythm", "_fix_rhythmic_discontinuities")}
    sig_c = inspect.signature(ref.PeakClassifier.classify_peaks)
    try:
        classifier.install(ref)
        corrections.install(ref)
        assert ref.PeakClassifier.classify_peaks is classifier.classify_peaks
        assert list(inspect.signature(ref.PeakClassifier.classify_peaks).parameters) == list(sig_c.parameters)
        for n, sig in sigs.items():
            assert list(inspect.signature(getattr(ref, n)).parameters) == list(sig.parameters), n
        got = stages_2_to_5()
    finally:
        importlib.reload(ref)
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]
    for i in (3, 4, 5):
        assert np.array_equal(got[i], want[i]) and got[i].dtype == want[i].dtype
    assert len(want[5]) > 100
    assert list(got[6]["beat_debug_info"].items()) == list(want[6]["beat_debug_info"].items())
    a, b = got[6]["long_term_bpm_series"], want[6]["long_term_bpm_series"]
    assert np.array_equal(a.values, b.values) and np.array_equal(a.index.values, b.index.values)
