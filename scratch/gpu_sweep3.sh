#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
run() { local label=$1; shift; local w=$1; shift; env "$@" $B --workload $w > gpurun_out/sw_$label.json 2> gpurun_out/sw_$label.err; echo "$label $w $(python -c "import json;l=json.load(open('gpurun_out/sw_$label.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])" 2>&1 | tail -1)"; }
run c4_d10 cfg4 A=1
run c4_d14 cfg4 MIA_RPPI2_DIV=14
run c4_d18 cfg4 MIA_RPPI2_DIV=18
run c2_d16 cfg2 MIA_RPPI2_DIV=16
run c2_d14r1 cfg2 MIA_RPPI2_DIV=14 MIA_RPPI2_RATIO=1
run small_d6 small A=1
run small_d10 small MIA_RPPI2_DIV=10
