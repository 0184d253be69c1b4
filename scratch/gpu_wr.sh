#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
run() { local label=$1; shift; local w=$1; shift; env "$@" $B --workload $w > gpurun_out/wr_$label.json 2> gpurun_out/wr_$label.err; echo "$label $w $(python -c "import json;l=json.load(open('gpurun_out/wr_$label.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])" 2>&1 | tail -1)"; }
run def cfg2 A=1
for v in w10 w7 w6; do run $v cfg2 MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_$v.so; done
run def_bins cfg2_default_bins MIA_RPPI_V2=2
run w10_bins cfg2_default_bins MIA_RPPI_V2=2 MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_w10.so
