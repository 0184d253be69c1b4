#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiled_kernels_match_general or full_size or reproducible or quarter" > gpurun_out/pytest_sym.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_sym.log
python bench.py --workload cfg2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/sym_cfg2.json 2> gpurun_out/sym_cfg2.err; echo "bench rc=$?"
cut -c1-900 gpurun_out/sym_cfg2.json
MIA_SYM=0 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-parity > gpurun_out/nosym_cfg2.json 2> gpurun_out/nosym_cfg2.err
cut -c1-300 gpurun_out/nosym_cfg2.json
