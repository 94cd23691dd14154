#!/bin/bash
mkdir -p gpurun_out
for lib in tools/tmp/libraster_static.so dynenv_b200/libdynenv_b200.so; do
  DYNENV_B200_LIB=$PWD/$lib python tools/gpu_raster_time.py 2>&1 | tail -1
done | tee gpurun_out/s2_raster.log
ncu --set full --clock-control none --import-source on -k regex:rc_phys_slow -s 600 -c 1 -o gpurun_out/prof_slow_r2 -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_slow_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obsp_kernelILb0 -s 1200 -c 1 -o gpurun_out/prof_obsp_r2 -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_obsp_r2.log 2>&1
tail -2 gpurun_out/ncu_obsp_r2.log
