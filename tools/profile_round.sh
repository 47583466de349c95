#!/bin/bash
# Full evidence refresh on the GPU box (one gpurun call): GPU test suite, smoke, bench line (+ reference arm), ncu launch list of
# the bench command, one `ncu --set full` capture of a cold MPC step (kernel table + hot lines of the five main kernels).
# usage: tools/profile_round.sh TAG  -> gpurun_out/TAG_*
TAG=${1:-r}
O=gpurun_out; mkdir -p $O /tmp/nc
(time timeout 400 python -m pytest tests -m gpu -q) > $O/${TAG}_tests.log 2>&1; tail -4 $O/${TAG}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2>> $O/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_bench_under_ncu.log 2>&1
tools/profile_ncu.sh $TAG
python tools/launch_table.py $O/${TAG}_launches.csv > $O/${TAG}_launch_shares.txt 2>&1
python - <<PY
import json
d = json.load(open("$O/${TAG}_bench.json"))
print("solves/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["stage_ms_per_solve"], d["cpu_baseline"])
PY
