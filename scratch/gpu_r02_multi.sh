#!/bin/bash
# usage: gpu_r02_multi.sh N [extra bench args]   (run under gpurun --gpus N)
N=$1; shift
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
fi
python - <<PY
import json
l=[json.loads(x) for x in open('gpurun_out/bench_n$N.json') if x.startswith('{')][-1]
print('value %.4g ms %.2f e2e %s'%(l['value'], l['ms_per_step'], l.get('e2e',{}).get('value')))
print('per_rank', l['per_rank_ms'])
print('phases', l['phases_ms'])
for k,v in l.get('secondary',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('per_rank_ms'), v.get('error'))
PY
