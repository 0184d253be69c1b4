#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --workload cfg2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
for v in b c d e f g; do
  MIA_LIB_PATH=$PWD/measure_ia_b200/lib/var/lib_$v.so $B > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  echo "$v $(python -c "import json;l=json.load(open('gpurun_out/var_$v.json'));print(l['ms_per_step'], l['config']['kernel'])")"
done
for t in 8 16 64; do
  MIA_TASKS_PER_WARP=$t $B > gpurun_out/var_t$t.json 2> gpurun_out/var_t$t.err
  echo "tpw$t $(python -c "import json;l=json.load(open('gpurun_out/var_t$t.json'));print(l['ms_per_step'], l['config']['kernel'])")"
done
for d in 8 12; do
  MIA_RPPI2_DIV=$d $B > gpurun_out/var_d$d.json 2> gpurun_out/var_d$d.err
  echo "div$d $(python -c "import json;l=json.load(open('gpurun_out/var_d$d.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])")"
done
MIA_RPPI2_RATIO=1 $B > gpurun_out/var_r1.json 2> gpurun_out/var_r1.err
echo "ratio1 $(python -c "import json;l=json.load(open('gpurun_out/var_r1.json'));print(l['ms_per_step'], l['config']['kernel'], l['config']['candidates_tested_per_step'])")"
