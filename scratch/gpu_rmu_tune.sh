#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --workload cfg3 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
run() { local label=$1; shift; env "$@" $B > gpurun_out/rt_$label.json 2> gpurun_out/rt_$label.err; echo "$label $(python -c "import json;l=json.load(open('gpurun_out/rt_$label.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])")"; }
run def A=1
for v in ru1 ru3 rc48; do run $v MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_$v.so; done
run div5 MIA_RMU_DIV=5
run div7 MIA_RMU_DIV=7
run div8 MIA_RMU_DIV=8
run lmul2 MIA_RMU_LMUL=2
run lmul4 MIA_RMU_LMUL=4
run tpw4 MIA_TASKS_PER_WARP=4
run tpw16 MIA_TASKS_PER_WARP=16
run ratio1 MIA_RMU_RATIO=1
run ratio3 MIA_RMU_RATIO=3
