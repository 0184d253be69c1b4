#!/bin/bash
# final-state run on one B200: default bench line (both arms), ncu launch list, full captures of the two hot kernels
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['e2e'], l['roofline']['frac'])
for k,v in l['secondary'].items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','wall_s','error','kernel')})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg2.csv \
  python bench.py --workload cfg2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary --no-parity > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tiled_rppi2s -s 2 -c 1 -o gpurun_out/r02_sym_final \
  python bench.py --workload cfg2 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity > gpurun_out/ncu_sym.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tiled_rmu -s 2 -c 1 -o gpurun_out/r02_rmu_final \
  python bench.py --workload cfg3 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity > gpurun_out/ncu_rmu.log 2>&1
ls -la gpurun_out | tail -8
