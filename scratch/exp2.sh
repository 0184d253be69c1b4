#!/bin/bash
# RMU unroll variants + rp-pi alignment, one line each
run() { # label env... 
  label=$1; shift
  out=$(env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/exp2.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'])" || { echo "$label FAILED"; tail -3 gpurun_out/exp2.err; }
}
WL=cfg3
run base X=1
run u6 MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/exp_u6.so
run u8 MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/exp_u8.so
run h2 MIA_RMU_HSPLIT=2
run d5 MIA_RMU_DIV=5
run d8 MIA_RMU_DIV=8
run d8h2 MIA_RMU_DIV=8 MIA_RMU_HSPLIT=2
WL=cfg2
run align0 MIA_RPPI_ALIGN=0
run align1 MIA_RPPI_ALIGN=1
run align2 MIA_RPPI_ALIGN=2
