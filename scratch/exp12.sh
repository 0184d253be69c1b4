#!/bin/bash
run() { label=$1; shift
  out=$(env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/exp12.err | tail -1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$WL', 'ms', round(d['ms_per_step'],2), 'pairs/s', round(d['value']/1e9,1),'e9 tested', d['config']['candidates_tested_per_step'], d['config']['kernel'])" || { echo "$label FAILED"; tail -3 gpurun_out/exp12.err; }
}
WL=cfg2
L=/root/repo/measure_ia_b200/lib
run old_swp MIA_RPPI_V2=0 MIA_LIB_PATH=$L/exp_swp_ni.so
run v2_swp_ni_d10 MIA_RPPI_V2=1 MIA_RPPI2_DIV=10 MIA_LIB_PATH=$L/exp_swp_ni.so
run v2_swp_inl_d10 MIA_RPPI_V2=1 MIA_RPPI2_DIV=10 MIA_LIB_PATH=$L/exp_swp_inl.so
run v2_swp_inl_u2_d10 MIA_RPPI_V2=1 MIA_RPPI2_DIV=10 MIA_LIB_PATH=$L/exp_swp_inl_u2.so
run v2_inl_d10 MIA_RPPI_V2=1 MIA_RPPI2_DIV=10 MIA_LIB_PATH=$L/exp_inl.so
run v2_inl_d12 MIA_RPPI_V2=1 MIA_RPPI2_DIV=12 MIA_LIB_PATH=$L/exp_inl.so
run v2_inl_d10_r3 MIA_RPPI_V2=1 MIA_RPPI2_DIV=10 MIA_RPPI2_RATIO=3 MIA_LIB_PATH=$L/exp_inl.so
