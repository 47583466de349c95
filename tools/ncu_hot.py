#!/usr/bin/env python3
"""Hot source lines + stall-reason mix of ONE kernel in an .ncu-rep (needs -lineinfo / --import-source on).
Usage: ncu_hot.py report.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
hdr = None; fname = ""; lines = []; stalls = {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        si = hdr.index("Warp Stall Sampling (All Samples)"); ii = hdr.index("Instructions Executed")
        try:
            lines.append((float(r[si]), float(r[ii]), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            continue
        for k, name in enumerate(hdr):
            if name.startswith("stall_") and "Not Issued" not in name:
                try: stalls[name] = stalls.get(name, 0) + float(r[k])
                except ValueError: pass
ts = sum(l[0] for l in lines) or 1; ti = sum(l[1] for l in lines) or 1
print(f"kernel ~ {kre}: {ti:.0f} warp instructions, {ts:.0f} stall samples")
tot = sum(stalls.values()) or 1
print("stall mix:", {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]})
print("--- stall-sample share | instruction share | file:line ---")
for l in sorted(lines, reverse=True)[:topn]:
    print(f"{100 * l[0] / ts:6.2f}% {100 * l[1] / ti:6.2f}%  {l[2]}:{l[3]:<4d} {l[4]}")
