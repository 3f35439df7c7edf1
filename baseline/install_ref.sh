#!/bin/bash
# Recipe: place the UNMODIFIED reference modules the hot path's callers live in under
# baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box with the snapshot).
#
# The reference is a desktop script, not a package (no setup.py / pyproject.toml), so
# `pip install --target baseline/_ref /root/reference` has nothing to build:
#   ERROR: Directory '/root/reference' is not installable. Neither 'setup.py' nor 'pyproject.toml' found.
# The two modules below are everything `import bpm_analysis` needs besides numpy / scipy / pandas
# (plotly is imported at module top and stubbed by oracle/load_reference.py; pydub is optional).
# Used by: tests/test_install_gpu.py (reference orchestrator + GPU front end vs the untouched
# module) and `bench.py --impl reference` (the reference's own functions on the host cores).
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
mkdir -p "$DST"
cp "$SRC/bpm_analysis.py" "$SRC/config.py" "$DST/"
mkdir -p "$DST/samples"
cp "$SRC"/samples/vulpine_filtered_debug.wav "$DST/samples/" 2>/dev/null || true
python - "$SRC" "$DST" <<'PY'
import hashlib, json, sys
src, dst = sys.argv[1], sys.argv[2]
man = {f: hashlib.sha256(open(f"{dst}/{f}", "rb").read()).hexdigest() for f in ("bpm_analysis.py", "config.py")}
json.dump({"source": src, "sha256": man}, open(f"{dst}/MANIFEST.json", "w"), indent=1)
print("installed", man)
PY
