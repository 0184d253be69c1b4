#!/bin/bash
run() { label=$1; shift
  out=$(env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/exp10.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'])" || { echo "$label FAILED"; tail -3 gpurun_out/exp10.err; }
}
WL=cfg2
for u in u1 u2 u3; do
run ${u}_d8 MIA_RPPI_V2=1 MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/exp_$u.so
run ${u}_d10 MIA_RPPI_V2=1 MIA_RPPI2_DIV=10 MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/exp_$u.so
done
