#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:rc_phys -s 110 -c 1 -o gpurun_out/prof_phys_late -f python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
