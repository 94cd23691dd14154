#!/bin/bash
# usage: gpu_ncu.sh <kernel regex> <launch skip> <out name> [steps of warm-up]
mkdir -p gpurun_out
W=${4:-30}
python tools/gpu_rc_variant.py $((W+2)) > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f python tools/gpu_rc_variant.py $((W+2)) > gpurun_out/ncu_$3.log 2>&1
tail -3 gpurun_out/ncu_$3.log
