#!/bin/bash
WL=${WL:-cfg2}
for v in base "$@"; do
  if [ $v = base ]; then lib=/root/repo/measure_ia_b200/lib/libmia_b200.so; else lib=/root/repo/measure_ia_b200/lib/exp_$v.so; fi
  out=$(MIA_LIB_PATH=$lib timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/exp4.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'])" || { echo "$v FAILED"; tail -3 gpurun_out/exp4.err; }
done
