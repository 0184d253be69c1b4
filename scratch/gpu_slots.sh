#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
run() { local label=$1; shift; local w=$1; shift; env "$@" $B --workload $w > gpurun_out/sl_$label.json 2> gpurun_out/sl_$label.err; echo "$label $w $(python -c "import json;l=json.load(open('gpurun_out/sl_$label.json'));print(l['ms_per_step'], l['phases_ms'], l['config']['kernel'])" 2>&1 | tail -1)"; }
for m in 1 2 4; do run m$m cfg2 MIA_SLOT_MULT=$m; done
for m in 1 2 4; do run m${m}_t8 cfg2 MIA_SLOT_MULT=$m MIA_TASKS_PER_WARP=8; done
for m in 1 2 4; do run c3m$m cfg3 MIA_SLOT_MULT=$m; done
run bins cfg2_default_bins A=1
run bins6 cfg2_default_bins MIA_RPPI2_DIV=6
run bins8 cfg2_default_bins MIA_RPPI2_DIV=8
run bins10 cfg2_default_bins MIA_RPPI2_DIV=10
