#!/bin/bash
run() { label=$1; shift
  out=$(env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/exp7.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'])" || { echo "$label FAILED"; tail -3 gpurun_out/exp7.err; }
}
for WL in cfg2 cfg3; do
run tpw8 MIA_TASKS_PER_WARP=8
run tpw16 MIA_TASKS_PER_WARP=16
run tpw32 MIA_TASKS_PER_WARP=32
run tpw64 MIA_TASKS_PER_WARP=64
done
