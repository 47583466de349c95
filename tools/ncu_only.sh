TAG=r02q; O=gpurun_out; mkdir -p $O /tmp/nc
H1_PROF_WORKLOAD=bench timeout 900 ncu --set full --clock-control none --import-source on -c 18 \
    --kernel-name 'regex:k_rollout|k_linearize|k_cost_quadratics|k_backward|k_line_search|k_primal_factor_seq' \
    -o /tmp/nc/full -f python tools/prof_run.py 4096 > $O/${TAG}_ncu.log 2>&1
tail -3 $O/${TAG}_ncu.log
python tools/ncu_kernels.py /tmp/nc/full.ncu-rep > $O/${TAG}_ncu_top_kernels.txt 2>&1
python tools/ncu_kernel_metrics.py /tmp/nc/full.ncu-rep 4096 $O/${TAG}_kernel_metrics.json
for k in k_backward k_line_search_quad k_linearize_tangents k_linearize_finish k_cost_quadratics; do
  python tools/ncu_hot.py /tmp/nc/full.ncu-rep $k 25 > $O/${TAG}_hot_$k.txt 2>&1
done
ls -la /tmp/nc/full.ncu-rep; ncu -i /tmp/nc/full.ncu-rep --page source --csv --print-source sass --kernel-name regex:k_line_search_quad 2>/dev/null | head -c 3000000 > $O/${TAG}_sass_ls_quad.csv
