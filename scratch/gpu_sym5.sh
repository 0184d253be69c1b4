#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
run() { # label, env..., workload
  local label=$1; shift; local w=$1; shift
  env "$@" $B --workload $w > gpurun_out/s5_$label.json 2> gpurun_out/s5_$label.err
  echo "$label $w $(python -c "import json;l=json.load(open('gpurun_out/s5_$label.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])")"
}
run def_cfg2 cfg2 A=1
run b_cfg2 cfg2 MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_b.so
run b2_cfg2 cfg2 MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_b2.so
run f_cfg2 cfg2 MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_f.so
run ord_cfg2 cfg2 MIA_SYM=0
run def_cfg3 cfg3 A=1
run def_bins cfg2_default_bins A=1
run v2_bins cfg2_default_bins MIA_RPPI_V2=2
run b_v2_bins cfg2_default_bins MIA_RPPI_V2=2 MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_b.so
run def_cfg4 cfg4 A=1
run b_cfg4 cfg4 MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_b.so
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
