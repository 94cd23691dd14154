#!/bin/bash
mkdir -p gpurun_out
export SPARSE=1
ncu --set full --clock-control none --import-source on -k regex:rc_phys_kernel -s 602 -c 1 -o gpurun_out/prof_r2_phys -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_r2_phys.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obsp_kernel -s 2408 -c 2 -o gpurun_out/prof_r2_obs -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_r2_obs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_phys_slow -s 601 -c 1 -o gpurun_out/prof_r2_slow -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_r2_slow.log 2>&1
tail -2 gpurun_out/ncu_r2_slow.log
ls -la gpurun_out/prof_r2_*
