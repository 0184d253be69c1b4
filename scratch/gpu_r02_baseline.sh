#!/bin/bash
# round-2 baseline on one B200: timings of the round-1 kernels and ncu captures of the FINAL (unroll-2) builds
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_base.csv &
SMI=$!
for w in cfg2 cfg2_default_bins cfg3 cfg3_default_bins; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/base_$w.json 2> gpurun_out/base_$w.err
done
kill $SMI
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tiled_rppi2 -s 2 -c 1 -o gpurun_out/r02_base_rppi2 \
  python bench.py --workload cfg2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_rppi2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tiled_rmu -s 2 -c 1 -o gpurun_out/r02_base_rmu \
  python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_rmu.log 2>&1
ls -la gpurun_out
