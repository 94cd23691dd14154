#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_phys -s 12 -c 1 -o gpurun_out/prof_phys -f python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_phys.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obs -s 12 -c 1 -o gpurun_out/prof_obs -f python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_obs.log 2>&1
tail -1 gpurun_out/ncu_obs.log
