#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_image.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -2
python bench.py --workload image --no-other > gpurun_out/bench_image.json 2> gpurun_out/bench_image.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_image.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['achieved'], d['roofline']['frac'], d['config']['episode_steps'])
PY
python tools/gpu_raster_time.py 23 > gpurun_out/raster_time.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_raster -s 2 -c 1 -o gpurun_out/prof_rc_raster -f python tools/gpu_raster_time.py 23 > gpurun_out/ncu_f7.log 2>&1
tail -1 gpurun_out/raster_time.log
python bench.py --workload image --envs 2048 --steps 2 --warmup 2 --no-cpu --no-other > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_image.csv python bench.py --workload image --envs 2048 --steps 2 --warmup 2 --no-cpu --no-other > gpurun_out/ncu_l3.log 2>&1
