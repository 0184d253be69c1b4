for v in exp_a exp_b exp_c exp_d exp_e exp_f; do
  MIA_LIB_PATH=/root/repo/measure_ia_b200/lib/$v.so timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,1), round(d['ms_per_step'],1), d['config']['candidates_tested_per_step'], d['config']['pairs_per_step'])"
done
