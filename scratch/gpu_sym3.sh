#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
for w in cfg2 cfg2_default_bins small; do
  $B --workload $w > gpurun_out/s3_$w.json 2> gpurun_out/s3_$w.err
  echo "$w $(python -c "import json;l=json.load(open('gpurun_out/s3_$w.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])")"
done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiled_kernels_match_general or full_size or reproducible or quarter" > gpurun_out/pytest_sym.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_sym.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tiled_rppi2s -s 2 -c 1 -o gpurun_out/r02_sym_v2 \
  python bench.py --workload cfg2 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity > gpurun_out/ncu_sym.log 2>&1
