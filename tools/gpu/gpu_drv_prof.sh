#!/bin/bash
# GPU box: Driving bench + ncu captures of its two kernels; quick RoboCup bench as a regression check
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 5 --no-cpu > gpurun_out/bench_rc_quick.json 2> gpurun_out/bench_rc_quick.err; tail -c 400 gpurun_out/bench_rc_quick.json
python bench.py --workload driving --steps 20 --warmup 10 > gpurun_out/bench_drv.json 2> gpurun_out/bench_drv.err; tail -c 2500 gpurun_out/bench_drv.json; tail -3 gpurun_out/bench_drv.err
python bench.py --workload driving --steps 3 --warmup 12 --no-cpu > gpurun_out/plain_drv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:drv_phys -s 12 -c 1 -o gpurun_out/prof_drv_phys -f python bench.py --workload driving --steps 3 --warmup 12 --no-cpu > gpurun_out/ncu_drv1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:drv_obs -s 12 -c 1 -o gpurun_out/prof_drv_obs -f python bench.py --workload driving --steps 3 --warmup 12 --no-cpu > gpurun_out/ncu_drv2.log 2>&1
tail -2 gpurun_out/ncu_drv2.log
