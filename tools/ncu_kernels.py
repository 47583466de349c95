#!/usr/bin/env python3
"""Per-kernel table of an .ncu-rep (ncu --set full): duration, occupancy, issue / fp64 / DMMA pipe activity, executed
fp64 operations (2 per DFMA, 1 per DADD / DMUL; DMMA m8n8k4 = 512 per instruction), DRAM traffic.
Usage: ncu_kernels.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
col = {h: i for i, h in enumerate(H)}
def g(r, name, default=0.0):
    try: return float(r[col[name]].replace(",", ""))
    except Exception: return default
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    ms = g(r, "gpu__time_duration.sum")
    unit = U[col["gpu__time_duration.sum"]]
    ms = ms / 1e3 if unit == "us" else (ms / 1e6 if unit == "ns" else ms)
    cyc = g(r, "sm__cycles_elapsed.max") or g(r, "smsp__cycles_elapsed.max")
    per_cycle = lambda op: g(r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")
    flop_cycle = 2 * per_cycle("dfma") + per_cycle("dadd") + per_cycle("dmul")
    dmma_pct = g(r, "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active")
    gb = lambda n: g(r, n) * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(U[col[n]], 1.0)
    print(f"{name}")
    print(f"  duration {ms:.3f} ms | grid {r[col['launch__grid_size']]} x {r[col['launch__block_size']]} | regs {r[col['launch__registers_per_thread']]}"
          f" | warps active {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % | issue active {g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} %")
    print(f"  fp64 pipe {g(r, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} % | DMMA pipe {dmma_pct:.1f} %"
          f" | executed fp64 (FMA pipe) {flop_cycle:.0f} flop/cycle of 18944 ({100 * flop_cycle / 18944:.1f} %)"
          f" = {flop_cycle * cyc / 1e9:.2f} GFLOP this launch")
    print(f"  dram read {gb('dram__bytes_read.sum'):.3f} GB, write {gb('dram__bytes_write.sum'):.3f} GB"
          f" | L1 hit {g(r, 'l1tex__t_sector_hit_rate.pct'):.0f} % | L2 hit {g(r, 'lts__t_sector_hit_rate.pct'):.0f} %"
          f" | warp instructions {g(r, 'smsp__inst_executed.sum'):.3e}")
