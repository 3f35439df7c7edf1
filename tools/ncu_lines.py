#!/usr/bin/env python
"""Attribute an ncu report's executed instructions and stall samples to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> [kernel-substring] [top_n]

Uses the source correlation stored in the report itself (`ncu --import-source on` at capture,
`ncu --page source --csv --print-source cuda,sass` here), so it does not depend on the library
that happens to be built in the tree.  One table per profiled launch whose name contains the
substring: share of warp instructions, share of stall samples, the top stall reasons of the line.
"""
import collections
import csv
import io
import subprocess
import sys

STALLS = ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg",
          "stall_membar", "stall_not_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst")


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    path = func = hdr = None
    launches = []                       # (function name, {(file, line): [inst, samples, text, stall counter]})
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            if func != r[1] or hdr is None or not launches:
                pass
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            if not launches or launches[-1][0] != func or path == launches[-1][2]:
                launches.append((func, collections.OrderedDict(), path))
            continue
        if hdr is None or func is None or r[0] == "":
            continue
        try:
            inst, smp = int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]])
        except (ValueError, KeyError):
            continue
        rec = launches[-1][1].setdefault((path, int(r[0])), [0, 0, r[1].strip()[:100], collections.Counter()])
        rec[0] += inst
        rec[1] += smp
        for s in STALLS:
            if s in hdr:
                try:
                    rec[3][s] += int(r[hdr[s]])
                except ValueError:
                    pass
    # merge the per-file sections of one launch (ncu prints one section per source file)
    merged = []
    for fn, lines, _ in launches:
        if merged and merged[-1][0] == fn and not (set(lines) & set(merged[-1][1])):
            merged[-1][1].update(lines)
        else:
            merged.append((fn, collections.OrderedDict(lines)))
    for fn, lines in merged:
        if want not in fn:
            continue
        ti = sum(v[0] for v in lines.values()) or 1
        ts = sum(v[1] for v in lines.values()) or 1
        print(f"== {fn}\n   warp instructions {ti}   stall samples {ts}")
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            why = ", ".join(f"{k[6:]} {100 * c / max(v[1], 1):.0f}%" for k, c in v[3].most_common(2) if c)
            print(f"   {f}:{ln:<5d} inst {100 * v[0] / ti:5.1f}%  smp {100 * v[1] / ts:5.1f}%  [{why}]  {v[2]}")
        print()


if __name__ == "__main__":
    main()
