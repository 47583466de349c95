#!/bin/bash
# `ncu --set full` capture of one cold MPC step (4096 instances, bench workload): first 14 launches of the solver's main kernels;
# kernel table + hot lines. usage: tools/profile_ncu.sh TAG -> gpurun_out/TAG_kernels.txt, TAG_hot_*.txt
TAG=${1:-r}
O=gpurun_out; mkdir -p $O /tmp/nc
H1_PROF_WORKLOAD=bench timeout 600 ncu --set full --clock-control none --import-source on -c 14 \
    --kernel-name 'regex:k_rollout_seq|k_linearize|k_cost_quadratics|k_backward|k_line_search_seq|k_primal_factor_seq' \
    -o /tmp/nc/full -f python tools/prof_run.py 4096 > $O/${TAG}_ncu.log 2>&1
python tools/ncu_kernels.py /tmp/nc/full.ncu-rep > $O/${TAG}_kernels.txt 2>&1
for k in k_backward k_line_search_seq k_linearize_tangents k_linearize_finish k_cost_quadratics; do
  python tools/ncu_hot.py /tmp/nc/full.ncu-rep $k 25 > $O/${TAG}_hot_$k.txt 2>&1
done
grep -A3 "^k_backward" $O/${TAG}_kernels.txt | head -4
