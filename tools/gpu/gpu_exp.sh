#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_rc_quiet.py 200 && ncu --set full --clock-control none --import-source on -k regex:rc_phys -s 1002 -c 1 -o gpurun_out/prof_quiet_late -f python tools/gpu_rc_quiet.py 200 > gpurun_out/ncu_q1.log 2>&1
python tools/gpu_rc_quiet.py 12 && ncu --set full --clock-control none --import-source on -k regex:rc_phys -s 62 -c 1 -o gpurun_out/prof_quiet_early -f python tools/gpu_rc_quiet.py 12 > gpurun_out/ncu_q2.log 2>&1
tail -1 gpurun_out/ncu_q2.log
