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
