#!/bin/bash
# One GPU-box round: the whole GPU test suite (or a -k selection in $SEL), smoke(), and a bench line.
# usage: tools/gpu_check.sh TAG [bench args]  -> gpurun_out/TAG_tests.log, TAG_smoke.log, TAG_bench.json
TAG=${1:-x}; shift
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
(time timeout ${TEST_TIMEOUT:-900} python -m pytest tests -m gpu -q ${PYTEST_X:+-x} --durations=15 ${SEL:+-k "$SEL"}) > $O/${TAG}_tests.log 2>&1; tail -25 $O/${TAG}_tests.log
if [ -z "$NOSMOKE" ]; then timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -3 $O/${TAG}_smoke.log; fi
if [ -z "$NOBENCH" ]; then
timeout 600 python bench.py "$@" > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -3 $O/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench.json"))
    print("solves/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms_per_solve"], "single", d.get("single_instance_ms_per_mpc_step"))
except Exception as e:
    print("bench parse failed", e)
PY
fi
