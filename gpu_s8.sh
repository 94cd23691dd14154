#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload image --envs 4096 --steps 2 --warmup 3 --no-cpu --no-other > gpurun_out/s8_img.json 2> gpurun_out/s8_img.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s8_img.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms_per_step'), d['roofline'].get('achieved'), d['roofline'].get('frac'))
PY
DYNENV_B200_LIB=$PWD/dynenv_b200/libdynenv_b200.so python tools/gpu_raster_time.py 23
