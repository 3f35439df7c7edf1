"""bpm_analysis_b200 -- the data-parallel front end of pixeru/bpm_analysis on NVIDIA B200.

Hand-written sm_100a CUDA (``csrc/``, built into ``libbpm_b200.so``) behind a C
ABI (``include/bpm_b200.h``), with a Python mirror of the reference's function
signatures (``frontend``) so the reference's sequential S1/S2 classifier, GUI
and web app can call it as a drop-in (``frontend.install``).
"""
from .params import HOT_PATH_DEFAULTS, default_params  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # lazy: importing the package must work without torch/CUDA (CPU-side tooling, oracle tests)
    if name in ("frontend", "runtime", "design", "synth", "dist", "classifier", "corrections"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
