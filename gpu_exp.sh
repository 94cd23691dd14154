#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_phys -s 60 -c 1 -o gpurun_out/prof_phys -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obs -s 60 -c 1 -o gpurun_out/prof_obs -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
