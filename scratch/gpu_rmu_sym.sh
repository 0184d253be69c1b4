#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
for w in cfg3 cfg3_default_bins cfg4_multipoles; do
  $B --workload $w > gpurun_out/rs_$w.json 2> gpurun_out/rs_$w.err
  echo "$w $(python -c "import json;l=json.load(open('gpurun_out/rs_$w.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])")"
done
MIA_RMU_SYM=0 $B --workload cfg4_multipoles > gpurun_out/rs0_cfg4m.json 2> gpurun_out/rs0_cfg4m.err
echo "cfg4_multipoles ordered $(python -c "import json;l=json.load(open('gpurun_out/rs0_cfg4m.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])")"
for u in 1 3; do :; done
