#!/bin/bash
run() { label=$1; shift
  out=$(env "$@" timeout 400 python bench.py --workload $WL --steps ${ST:-3} --warmup ${WU:-2} --no-e2e --no-cpu-baseline 2>gpurun_out/exp15.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'], 'frac', round(d['roofline']['frac'],3))" || { echo "$label FAILED"; tail -3 gpurun_out/exp15.err; }
}
WL=cfg2
run new X=1
run new X=1
WL=small
run new X=1
WL=cfg2_default_bins
run new X=1
