#!/usr/bin/env python
"""Copy the evidence of one `tools/capture_evidence.sh <tag>` run from gpurun_out/ into profiles/
(tracked) and rebuild profiles/ncu_traffic.json (DRAM bytes per launch of each profiled kernel,
read by bench.py for roofline.traffic).

    python tools/collect_profiles.py <tag> [round]
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(REPO, "gpurun_out"), os.path.join(REPO, "profiles")


def traffic(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return {}
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").replace("bpm::", "").strip()
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[col[k]]) * scale.get(units[col[k]], 1.0)
        acc.setdefault(name, []).append(b)
    return {k: sum(v) / len(v) for k, v in acc.items()}


def main():
    tag = sys.argv[1]
    rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
    os.makedirs(P, exist_ok=True)
    for f in sorted(os.listdir(G)):
        if not f.startswith(f"ev_{tag}_"):
            continue
        if f.endswith((".ncu-rep", ".err")) or "plain" in f or "ncu_full" in f or "ncu_launches" in f:
            continue
        shutil.copy(os.path.join(G, f), os.path.join(P, f"{rnd}_{f[3:]}"))
        print("copied", f)
    tr = {}
    for mode in ("parity", "fullrate"):
        rep = os.path.join(G, f"ev_{tag}_full_{mode}.ncu-rep")
        if os.path.exists(rep):
            tr[mode] = traffic(rep)
    if tr:
        tr["_source"] = f"ncu --set full, tools/capture_evidence.sh {tag}: mean dram__bytes_read.sum + dram__bytes_write.sum per launch"
        with open(os.path.join(P, "ncu_traffic.json"), "w") as fh:
            json.dump(tr, fh, indent=1)
        print(json.dumps(tr, indent=1))


if __name__ == "__main__":
    main()
