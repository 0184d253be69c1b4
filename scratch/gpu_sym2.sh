#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --workload cfg2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
for v in c2 c2w5 u1 c2u1; do
  MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_$v.so $B > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  echo "$v $(python -c "import json;l=json.load(open('gpurun_out/var_$v.json'));print(l['ms_per_step'], l['config']['kernel'])")"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tiled_rppi2s -s 2 -c 1 -o gpurun_out/r02_sym_v1 \
  python bench.py --workload cfg2 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity > gpurun_out/ncu_sym.log 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiled_kernels_match_general or full_size or reproducible or quarter" > gpurun_out/pytest_sym.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_sym.log
