#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:rc_raster -s 2 -c 1 -o gpurun_out/prof_r2_raster -f python tools/gpu_raster_time.py 23 > gpurun_out/ncu_r2_raster.log 2>&1
tail -2 gpurun_out/ncu_r2_raster.log
