#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
run() { local label=$1; shift; local w=$1; shift; env "$@" $B --workload $w > gpurun_out/sw_$label.json 2> gpurun_out/sw_$label.err; echo "$label $w $(python -c "import json;l=json.load(open('gpurun_out/sw_$label.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])" 2>&1 | tail -1)"; }
run c4_d22 cfg4 MIA_RPPI2_DIV=22
run c4_d26 cfg4 MIA_RPPI2_DIV=26
run c4_d30 cfg4 MIA_RPPI2_DIV=30
run c2_d18 cfg2 MIA_RPPI2_DIV=18
run c2_d20 cfg2 MIA_RPPI2_DIV=20
