#!/bin/bash
# ncu --set full of the RoboCup step kernels at step $1 (default 100) of the bench workload
mkdir -p gpurun_out
W=${1:-100}
python tools/gpu_rc_variant.py $((W+2)) > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
for k in ${2:-rc_phys_kernel rc_phys_slow "rc_obsp_kernel<true>" "rc_obsp_kernel<false>"}; do
  case $k in rc_phys_kernel) n=5; o=prof_phys;; rc_phys_slow) n=5; o=prof_slow;; *true*) n=10; o=prof_obsl;; *) n=10; o=prof_obsp;; esac
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s $((W*n)) -c 1 -o gpurun_out/$o -f python tools/gpu_rc_variant.py $((W+2)) > gpurun_out/ncu_$o.log 2>&1
  tail -1 gpurun_out/ncu_$o.log
done
