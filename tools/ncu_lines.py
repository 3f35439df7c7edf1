#!/usr/bin/env python
"""Attribute an ncu report's per-instruction samples to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <mangled-kernel-substring> [top_n]

Uses `ncu --page source --csv` (SASS view with sampling / executed-instruction counts) and
`nvdisasm -g` line info of the cubin inside bpm_analysis_b200/libbpm_b200.so.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kname = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(REPO, "bpm_analysis_b200", "libbpm_b200.so")],
                   cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
    start = [i for i, l in enumerate(sass) if l.startswith("//") and ".text." in l and kname in l][0]
    end = [i for i, l in enumerate(sass) if l.startswith("//---------------------") and i > start]
    end = end[0] if end else len(sass)
    cur, off2line = None, {}
    for l in sass[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # the report may hold several kernels: pick the first block whose name matches
    blocks, curb = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            curb = {"name": r[1], "rows": []}
            blocks.append(curb)
        elif curb is not None:
            curb["rows"].append(r)
    blk = [b for b in blocks if kname.split("EPK")[0].replace("_ZN3bpm", "")[2:] in b["name"] or kname in b["name"]]
    blk = blk[0] if blk else blocks[0]
    hdr = blk["rows"][0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg_s, agg_i, agg_t = collections.Counter(), collections.Counter(), collections.Counter()
    base, tot_s, tot_i = None, 0, 0
    for r in blk["rows"][1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[idx["Address"]], 16)
        base = addr if base is None else base
        key = off2line.get(addr - base)
        s, i_, t = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]]), int(r[idx["Thread Instructions Executed"]])
        agg_s[key] += s; agg_i[key] += i_; agg_t[key] += t
        tot_s += s; tot_i += i_
    print(f"kernel: {blk['name'][:80]}\ntotal samples {tot_s}  warp instructions {tot_i}")
    cache = {}
    for key, s in agg_s.most_common(top):
        txt = ""
        if key:
            path = os.path.join(REPO, "bpm_analysis_b200", "csrc", key[0])
            if os.path.exists(path):
                cache.setdefault(path, open(path).read().splitlines())
                txt = cache[path][key[1] - 1].strip()[:88]
        print(f"{str(key):26s} smp {100 * s / max(tot_s, 1):5.1f}%  inst {100 * agg_i[key] / max(tot_i, 1):5.1f}% "
              f"({agg_i[key]:>10d})  lanes {agg_t[key] / max(1, agg_i[key]):4.1f} | {txt}")


if __name__ == "__main__":
    main()
