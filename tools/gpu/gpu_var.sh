#!/bin/bash
# quick: RoboCup + Driving timing over N steps for every library variant given (default: the in-tree build)
mkdir -p gpurun_out
N=${N:-100}
for lib in ${@:-dynenv_b200/libdynenv_b200.so}; do
  DYNENV_B200_LIB=$PWD/$lib timeout 300 python tools/gpu_rc_variant.py $N 2>&1 | tail -1
done | tee gpurun_out/var.log
