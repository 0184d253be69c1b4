#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
run() { local label=$1; shift; local w=$1; shift; env "$@" $B --workload $w > gpurun_out/sw_$label.json 2> gpurun_out/sw_$label.err; echo "$label $w $(python -c "import json;l=json.load(open('gpurun_out/sw_$label.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])" 2>&1 | tail -1)"; }
run base cfg2 A=1
run ratio1 cfg2 MIA_RPPI2_RATIO=1
run ratio1d8 cfg2 MIA_RPPI2_RATIO=1 MIA_RPPI2_DIV=8
run div8 cfg2 MIA_RPPI2_DIV=8
run div12 cfg2 MIA_RPPI2_DIV=12
run div14 cfg2 MIA_RPPI2_DIV=14
run ratio3 cfg2 MIA_RPPI2_RATIO=3 MIA_RPPI2_DIV=12
run t4 cfg2 MIA_TASKS_PER_WARP=4
run bins_d12 cfg2_default_bins MIA_RPPI2_DIV=12
run bins_r1 cfg2_default_bins MIA_RPPI2_RATIO=1
