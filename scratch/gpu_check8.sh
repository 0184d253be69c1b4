#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
run() { local label=$1; shift; local w=$1; shift; env "$@" $B --workload $w > gpurun_out/ck_$label.json 2> gpurun_out/ck_$label.err; echo "$label $w $(python -c "import json;l=json.load(open('gpurun_out/ck_$label.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])" 2>&1 | tail -1)"; }
run cfg2 cfg2 A=1
run cfg2o cfg2 MIA_SYM=0
run cfg4 cfg4 A=1
run bins cfg2_default_bins A=1
run small small A=1
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
