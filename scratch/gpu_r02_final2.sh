#!/bin/bash
# final-state run of round 2 (after the light-cone work) on one B200: full GPU test suite, smoke(), default bench line,
# ncu launch list + full capture of the light-cone kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final2.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final2.log | cut -c1-200
timeout 400 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l.get('parity_check',{}).get('dd_and_dd_jk_bit_exact'))
for k,v in l['secondary'].items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','wall_s','error','kernel','batched','pair_kernel_ms_rank0','raw_pairs_per_s')})
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_lightcone -s 1 -c 1 -o gpurun_out/r02_lightcone \
  python scratch/lc_profile.py > gpurun_out/ncu_lc.log 2>&1; echo "ncu rc=$?"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_lightcone.csv \
  python scratch/lc_profile.py > gpurun_out/launches_lc.log 2>&1
ls -la gpurun_out | tail -6
