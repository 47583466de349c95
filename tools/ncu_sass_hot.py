#!/usr/bin/env python3
"""Hottest SASS instructions (by warp stall samples) of ONE kernel in an .ncu-rep. Usage: ncu_sass_hot.py rep kernel_regex [n]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; lines = []
for r in rows:
    if not r: continue
    if "Source" in r and "Warp Stall Sampling (All Samples)" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        try:
            lines.append((float(r[hdr.index("Warp Stall Sampling (All Samples)")]), float(r[hdr.index("Instructions Executed")]), r[hdr.index("Address")] if "Address" in hdr else "", r[hdr.index("Source")], r))
        except ValueError:
            pass
ts = sum(l[0] for l in lines) or 1
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_")] if hdr else []
for k, l in enumerate(lines):
    pass
order = sorted(range(len(lines)), key=lambda i: -lines[i][0])[:topn]
for i in order:
    l = lines[i]
    top = sorted(((float(l[4][c]) if l[4][c] else 0.0, hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{100 * l[0] / ts:6.2f}%  idx {i:5d}  exec {l[1]:>10.0f}  {l[3][:90]:90s}  {top}")
    # context: previous 2 instructions
