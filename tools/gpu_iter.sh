#!/bin/bash
# One measurement round on the GPU box: Riccati / solve parity subset (or the whole GPU suite with FULL=1) + a bench line without the CPU leg.
# usage: tools/gpu_iter.sh TAG   -> gpurun_out/TAG_tests.log, gpurun_out/TAG_bench.json
TAG=${1:-x}
mkdir -p gpurun_out
if [ -n "$FULL" ]; then SEL=""; else SEL="-k backward or solve_parity or batch_equals or golden or full_size"; fi
timeout 400 python -m pytest tests -m gpu -q -x ${SEL:+"$SEL"} > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("solves/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms_per_solve"], "single", d["single_instance_ms_per_mpc_step"])
PY
