#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_robocup.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3 )
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -3 gpurun_out/bench_r2c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2c.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['roofline']['share_of_step'])
print({k:(v.get('value'),v.get('e2e')) for k,v in d['other_workloads'].items()})
print(d.get('arranger'))
PY
