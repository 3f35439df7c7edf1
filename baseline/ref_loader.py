"""Loads the UNMODIFIED reference ``bpm_analysis`` module from ``baseline/_ref`` (placed there by
``baseline/install_ref.sh``; git-ignored, travels to the GPU box).  Used by ``bench.py --impl
reference`` and by the install tests; the product package never imports this.

The reference imports plotly at module top (``bpm_analysis.py:7-8``) although only ``Plotter``
uses it; plotly is not installed in this image, so empty stub modules are injected first.  Nothing
else is altered.  ``fresh=True`` returns a new, independent module object (so that one copy can be
re-bound by ``frontend.install`` while another stays untouched)."""
from __future__ import annotations

import importlib.util
import logging
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "bpm_analysis.py"))


def _stub_plotly() -> None:
    if "plotly" in sys.modules:
        return
    try:
        import plotly  # noqa: F401
        return
    except ImportError:
        pass
    plotly = types.ModuleType("plotly")
    go = types.ModuleType("plotly.graph_objects")
    sub = types.ModuleType("plotly.subplots")
    sub.make_subplots = lambda *a, **k: None
    plotly.graph_objects, plotly.subplots = go, sub
    sys.modules.update({"plotly": plotly, "plotly.graph_objects": go, "plotly.subplots": sub})


def load(fresh: bool = False, name: str = "bpm_analysis_ref"):
    if not available():
        raise FileNotFoundError(f"{REF_DIR}/bpm_analysis.py missing: run baseline/install_ref.sh")
    _stub_plotly()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)           # the module does `from config import DEFAULT_PARAMS`-style imports
    level = logging.getLogger().level
    if not fresh and name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name if not fresh else f"{name}_{len(sys.modules)}",
                                                  os.path.join(REF_DIR, "bpm_analysis.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    logging.getLogger().setLevel(level if level else logging.WARNING)     # the module sets INFO at import
    if not fresh:
        sys.modules[name] = mod
    return mod


def default_params():
    spec = importlib.util.spec_from_file_location("bpm_config_ref", os.path.join(REF_DIR, "config.py"))
    cfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfg)
    return dict(cfg.DEFAULT_PARAMS)
