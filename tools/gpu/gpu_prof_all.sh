#!/bin/bash
# ncu --set full of the RoboCup step kernels at step $1 (default 100) of the bench workload
mkdir -p gpurun_out
W=${1:-100}
python tools/gpu_rc_variant.py $((W+2)) > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
for o in ${2:-prof_phys prof_slow prof_obsl prof_obsp}; do
  case $o in prof_phys) k=rc_phys_kernel; s=$((W*5));; prof_slow) k=rc_phys_slow; s=$((W*5));; prof_obsl) k=rc_obsp_kernel; s=$((W*20));; *) k=rc_obsp_kernel; s=$((W*20+1));; esac
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s $s -c 1 -o gpurun_out/$o -f python tools/gpu_rc_variant.py $((W+2)) > gpurun_out/ncu_$o.log 2>&1
  tail -1 gpurun_out/ncu_$o.log
done
