#!/usr/bin/env python3
"""Writes profiles/<tag>_kernel_metrics.json from an `ncu --set full` report of tools/prof_run.py (B instances, bench
workload, first launches of a cold MPC step): per-knot DRAM traffic of the Riccati / linearization / cost kernels and the
executed fp64 operations per f_D evaluation of the line-search kernel. bench.py reads these figures instead of carrying
literals. Usage: ncu_kernel_metrics.py report.ncu-rep B out.json"""
import csv, io, json, subprocess, sys
rep, B, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
N = 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
col = {h: i for i, h in enumerate(H)}
def g(r, name):
    try: return float(r[col[name]].replace(",", ""))
    except Exception: return 0.0
def gb(r, n): return g(r, n) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(U[col[n]], 1.0)
first = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("h1::", "")
    key = name.split("<")[0]
    grid = g(r, "launch__grid_size")
    if key in first and first[key]["grid"] >= grid:   # keep the largest (full-batch) launch of each kernel
        if not key.startswith("k_linearize_tangents"): continue
    cyc = g(r, "sm__cycles_elapsed.max") or g(r, "smsp__cycles_elapsed.max")
    pc = lambda op: g(r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")
    fl = (2 * pc("dfma") + pc("dadd") + pc("dmul")) * cyc
    unit = U[col["gpu__time_duration.sum"]]
    ms = g(r, "gpu__time_duration.sum"); ms = ms / 1e3 if unit == "us" else (ms / 1e6 if unit == "ns" else ms)
    e = {"grid": grid, "ms_under_ncu": ms, "dram_bytes": gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum"),
         "fp64_fma_pipe_flops": fl, "regs": g(r, "launch__registers_per_thread"),
         "dmma_pipe_pct": g(r, "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
         "fp64_pipe_pct": g(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
         "warps_active_pct": g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
         "local_store_bytes": gb(r, "smsp__inst_executed_op_local_st.sum") if "smsp__inst_executed_op_local_st.sum" in col else None}
    if key.startswith("k_linearize_tangents") and key in first:
        for k in ("dram_bytes", "fp64_fma_pipe_flops", "ms_under_ncu"): first[key][k] += e[k]
    else:
        first[key] = e
knots = B * N
res = {"source": rep.split("/")[-1], "instances": B, "kernels": first}
if "k_backward" in first: res["backward_dram_bytes_per_knot"] = first["k_backward"]["dram_bytes"] / knots
lin = [first[k] for k in first if k.startswith("k_linearize")]
if lin: res["linearize_dram_bytes_per_knot"] = sum(k["dram_bytes"] for k in lin) / knots
if "k_cost_quadratics" in first: res["cost_quadratics_dram_bytes_per_knot"] = first["k_cost_quadratics"]["dram_bytes"] / (B * (N + 1))
for k in first:
    if k.startswith("k_line_search"):
        res["line_search_flops_per_eval"] = first[k]["fp64_fma_pipe_flops"] / (B * 8 * N)
        res["line_search_dram_bytes_per_instance"] = first[k]["dram_bytes"] / B
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "kernels"}))
