#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/s11_pytest.log 2>&1; cat gpurun_out/s11_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -3 gpurun_out/bench_r2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['config'].get('obs_stream'))
print({k:(v.get('value'),v.get('e2e')) for k,v in d['other_workloads'].items()})
print(d.get('arranger'))
PY
