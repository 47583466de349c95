#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics + the hottest source lines (needs -lineinfo). Usage: ncu_summary.py rep [n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, V = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    for i, h in enumerate(H):
        if h == w:
            print(f"{h:70s} {V[i]:>22s} {U[i]}")
for i, h in enumerate(H):
    if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
        try:
            if float(V[i]) > 3.0: print(f"{h:70s} {V[i]:>22s} %")
        except ValueError: pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# rows with a line number carry the per-CUDA-line aggregate
hdr = [k for k, r in enumerate(rows) if r and r[0] == "Line No"]
if hdr:
    Hs = rows[hdr[0]]
    si = Hs.index("Warp Stall Sampling (All Samples)")
    ii = Hs.index("Instructions Executed")
    lines = []
    fname = ""
    for r in rows:
        if r and r[0] == "File Name": fname = r[1].split("/")[-1]
        if r and r[0].isdigit() and len(r) > ii:
            try: lines.append((float(r[si]), float(r[ii]), fname, int(r[0]), r[1].strip()[:100]))
            except ValueError: pass
    ts = sum(l[0] for l in lines) or 1; ti = sum(l[1] for l in lines) or 1
    print("--- hottest CUDA source lines: stall-sample share | instruction share | file:line ---")
    for l in sorted(lines, reverse=True)[:topn]:
        print(f"{100 * l[0] / ts:6.2f}% {100 * l[1] / ti:6.2f}%  {l[2]}:{l[3]:<4d} {l[4]}")
