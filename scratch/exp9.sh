#!/bin/bash
run() { label=$1; shift
  out=$(env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/exp9.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'])" || { echo "$label FAILED"; tail -3 gpurun_out/exp9.err; }
}
WL=cfg2
run old MIA_RPPI_V2=0
run d4r1 MIA_RPPI_V2=1 MIA_RPPI2_DIV=4 MIA_RPPI2_RATIO=1
run d5r1 MIA_RPPI_V2=1 MIA_RPPI2_DIV=5 MIA_RPPI2_RATIO=1
run d6r1 MIA_RPPI_V2=1 MIA_RPPI2_DIV=6 MIA_RPPI2_RATIO=1
run d6r2 MIA_RPPI_V2=1 MIA_RPPI2_DIV=6 MIA_RPPI2_RATIO=2
run d8r2 MIA_RPPI_V2=1 MIA_RPPI2_DIV=8 MIA_RPPI2_RATIO=2
run d10r2 MIA_RPPI_V2=1 MIA_RPPI2_DIV=10 MIA_RPPI2_RATIO=2
run d12r3 MIA_RPPI_V2=1 MIA_RPPI2_DIV=12 MIA_RPPI2_RATIO=3
run d12r2 MIA_RPPI_V2=1 MIA_RPPI2_DIV=12 MIA_RPPI2_RATIO=2
run d8r2t8 MIA_RPPI_V2=1 MIA_TASKS_PER_WARP=8
