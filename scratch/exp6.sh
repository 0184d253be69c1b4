#!/bin/bash
run() { label=$1; shift
  out=$(env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/exp6.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'])" || { echo "$label FAILED"; tail -3 gpurun_out/exp6.err; }
}
WL=cfg3
run base X=1
run u2 MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/exp_ru2.so
run u3 MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/exp_ru3.so
run ch96 MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/exp_rch96.so
run ch48 MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/exp_rch48.so
run d7 MIA_RMU_DIV=7
run d7r3 MIA_RMU_DIV=7 MIA_RMU_RATIO=3
run d6r2u3 MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/exp_ru3.so MIA_RMU_HSPLIT=2
