import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not errored) on a host without a CUDA device or without the
    native library: a plain `pytest` on a CPU box runs the CPU suite only."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:                                           # noqa: BLE001
        have_gpu = False
    have_lib = os.path.exists(os.path.join(REPO, "bpm_analysis_b200", "libbpm_b200.so"))
    if have_gpu and have_lib:
        return
    why = "no CUDA device" if not have_gpu else "bpm_analysis_b200/libbpm_b200.so not built"
    skip = pytest.mark.skip(reason=why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def ref_params():
    """Hot-path params at the reference's defaults (config.py:3-108), debug WAV off."""
    from bpm_analysis_b200.params import default_params
    p = default_params()
    p["save_filtered_wav"] = False
    return p


@pytest.fixture(scope="session")
def synth_inputs():
    """name -> (pcm, sample_rate); identical to what oracle/make_golden.py fed the reference."""
    from oracle.make_golden import synth_cases
    return synth_cases()


def rel_err(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY.md §8(c)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "NaN positions differ"
    scale = np.max(np.abs(b[~nan_b])) if np.any(~nan_b) else 1.0
    if scale == 0:
        scale = 1.0
    return float(np.max(np.abs(a[~nan_a] - b[~nan_b])) / scale) if np.any(~nan_a) else 0.0
