#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 200 --no-cpu > gpurun_out/plain_late.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_phys -s 1002 -c 1 -o gpurun_out/prof_phys_late -f python bench.py --steps 2 --warmup 200 --no-cpu > gpurun_out/ncu_late.log 2>&1
tail -2 gpurun_out/ncu_late.log
