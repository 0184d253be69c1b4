#!/bin/bash
# usage: rmu_sweep.sh "D R H" ...
for cfg in "$@"; do
  set -- $cfg
  out=$(MIA_RMU_DIV=$1 MIA_RMU_RATIO=$2 MIA_RMU_HSPLIT=$3 timeout 200 python bench.py --workload ${WL:-cfg3} --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/sweep.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('div $1 ratio $2 hsplit $3', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], 'ratio', round(d['config']['candidates_tested_per_step']/d['config']['pairs_per_step'],2), d['config']['kernel'])" || { echo "div $1 ratio $2 hsplit $3 FAILED"; tail -3 gpurun_out/sweep.err; }
done
