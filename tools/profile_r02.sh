#!/bin/bash
# Evidence refresh of round 2 on the GPU box (one gpurun call): GPU test suite, smoke, bench line (+ reference arm), ncu launch
# list of the bench command, one `ncu --set full` capture of a cold MPC step at 4096 instances (kernel table, hot lines,
# kernel metrics JSON for bench.py). usage: tools/profile_r02.sh TAG -> gpurun_out/TAG_*
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O /tmp/nc
(time timeout 900 python -m pytest tests -m gpu -q) > $O/${TAG}_tests.log 2>&1; tail -4 $O/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
H1_PROF_WORKLOAD=bench timeout 900 ncu --set full --clock-control none --import-source on -c 18 \
    --kernel-name 'regex:k_rollout|k_linearize|k_cost_quadratics|k_backward|k_line_search|k_primal_factor_seq' \
    -o /tmp/nc/full -f python tools/prof_run.py 4096 > $O/${TAG}_ncu.log 2>&1
python tools/ncu_kernels.py /tmp/nc/full.ncu-rep > $O/${TAG}_ncu_top_kernels.txt 2>&1
python tools/ncu_kernel_metrics.py /tmp/nc/full.ncu-rep 4096 $O/${TAG}_kernel_metrics.json
for k in k_backward k_line_search_quad k_linearize_tangents k_linearize_finish k_cost_quadratics; do
  python tools/ncu_hot.py /tmp/nc/full.ncu-rep $k 25 > $O/${TAG}_hot_$k.txt 2>&1
done
cp $O/${TAG}_kernel_metrics.json profiles/r02_kernel_metrics.json
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -2 $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2>> $O/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/${TAG}_bench_under_ncu.log 2>&1
python tools/launch_table.py $O/${TAG}_launches.csv > $O/${TAG}_launch_shares.txt 2>&1
python - <<PY
import json
d = json.load(open("$O/${TAG}_bench.json"))
print("solves/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["stage_ms_per_solve"])
print("warm", d["warm_closed_loop"]); print("single", d["single_instance"]); print("cpu", d["cpu_baseline"]); print("finite", d["all_finite"], d["instances_status_ok"])
print({k: (round(v["ms_per_launch"], 3), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
PY
