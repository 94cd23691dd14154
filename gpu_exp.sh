#!/bin/bash
# GPU box: ncu capture of the Driving physics kernel late in the warm-up (contacts present)
mkdir -p gpurun_out
python bench.py --workload driving --steps 3 --warmup 80 --no-cpu > gpurun_out/plain_drv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:drv_phys -s 80 -c 1 -o gpurun_out/prof_drv_phys_late -f python bench.py --workload driving --steps 3 --warmup 80 --no-cpu > gpurun_out/ncu_drv1.log 2>&1
tail -2 gpurun_out/ncu_drv1.log
