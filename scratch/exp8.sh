#!/bin/bash
run() { label=$1; shift
  out=$(env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/exp8.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'], d['phases_ms'])" || { echo "$label FAILED"; tail -3 gpurun_out/exp8.err; }
}
WL=cfg3
run l1 MIA_RMU_LMUL=1
run l2 MIA_RMU_LMUL=2
run l3 MIA_RMU_LMUL=3
run l4 MIA_RMU_LMUL=4
run l2d5 MIA_RMU_LMUL=2 MIA_RMU_DIV=5
run l3d5 MIA_RMU_LMUL=3 MIA_RMU_DIV=5
run l3d4r1 MIA_RMU_LMUL=3 MIA_RMU_DIV=4 MIA_RMU_RATIO=1
