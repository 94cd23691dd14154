#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('env-steps/s', d['value'], 'step_ms', d['ms_per_step'], d['roofline']['launch_ms'], 'e2e', d['e2e']['value'])"
if [ "$1" == "prof" ]; then
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_step -s 3 -c 1 -o gpurun_out/prof -f python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
fi
