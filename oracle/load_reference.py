"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (authoring container only).

Imports the UNMODIFIED reference module from ``/root/reference`` so that
``oracle/make_golden.py`` can record its outputs as fixtures and
``tests/test_oracle_golden.py`` can (when the tree is present) re-check the
restatement live.  ``/root/reference`` does not exist on the GPU box, so nothing
marked ``gpu``, nor ``smoke()``, nor ``bench.py`` may call this.

The reference imports plotly at module top (``bpm_analysis.py:7-8``) although
only its ``Plotter`` uses it; plotly is not installed here, so two empty stub
modules are injected first.  Nothing else is altered.
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BPM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "bpm_analysis.py"))


def load_reference():
    """Return the reference's ``bpm_analysis`` module (imported in-process)."""
    if not reference_available():
        raise FileNotFoundError(f"no reference tree at {REFERENCE_ROOT}")
    if "plotly" not in sys.modules:
        try:
            importlib.import_module("plotly")
        except ImportError:
            plotly = types.ModuleType("plotly")
            go = types.ModuleType("plotly.graph_objects")
            sub = types.ModuleType("plotly.subplots")
            sub.make_subplots = lambda *a, **k: None
            plotly.graph_objects, plotly.subplots = go, sub
            sys.modules.update({"plotly": plotly, "plotly.graph_objects": go,
                                "plotly.subplots": sub})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mod = importlib.import_module("bpm_analysis")
    logging.getLogger().setLevel(logging.WARNING)   # the module sets INFO at import
    return mod


def reference_params():
    """The reference's DEFAULT_PARAMS (``config.py:3-108``), a fresh copy."""
    load_reference()
    cfg = importlib.import_module("config")
    return dict(cfg.DEFAULT_PARAMS)
