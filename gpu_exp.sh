#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu --debug-full-obs 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('FULL env-steps/s', d['value'], 'kernel_ms', d['roofline']['kernel_ms'])"
ncu --set full --clock-control none --import-source on -k regex:rc_step -s 3 -c 1 -o gpurun_out/prof_full -f python bench.py --steps 3 --warmup 2 --no-cpu --debug-full-obs > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
