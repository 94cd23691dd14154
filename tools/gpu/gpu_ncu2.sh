#!/bin/bash
# usage: gpu_ncu2.sh <kernel regex> <launch skip> <out name> <steps of warm-up>
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f python tools/gpu_rc_variant.py $(($4+2)) > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/ncu_$3.log
