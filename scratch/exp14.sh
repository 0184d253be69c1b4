#!/bin/bash
run() { label=$1; shift
  out=$(env "$@" timeout 400 python bench.py --workload $WL --steps ${ST:-3} --warmup ${WU:-2} --no-e2e --no-cpu-baseline 2>gpurun_out/exp14.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'], 'frac', round(d['roofline']['frac'],3))" || { echo "$label FAILED"; tail -3 gpurun_out/exp14.err; }
}
WL=cfg2_default_bins
run d8 MIA_RPPI2_DIV=8
run d6 MIA_RPPI2_DIV=6
run d5 MIA_RPPI2_DIV=5
run d6r1 MIA_RPPI2_DIV=6 MIA_RPPI2_RATIO=1
run d4r1 MIA_RPPI2_DIV=4 MIA_RPPI2_RATIO=1
WL=small
run d6 MIA_RPPI2_DIV=6
run d4r1 MIA_RPPI2_DIV=4 MIA_RPPI2_RATIO=1
WL=cfg2
run d6 MIA_RPPI2_DIV=6
run d8 MIA_RPPI2_DIV=8
ST=1 WU=1 WL=cfg4
run d12 MIA_RPPI2_DIV=12
run d16 MIA_RPPI2_DIV=16
