#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}; EXTRA="${@:2}"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -2 gpurun_out/bench_${N}gpu.err; python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open('gpurun_out/bench_%sgpu.json'%n).read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['config']['envs_total'], d['clocks'])
PY
