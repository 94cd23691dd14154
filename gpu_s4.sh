#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_robocup.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s4_pytest.log 2>&1
tail -5 gpurun_out/s4_pytest.log
N=130 ./gpu_var.sh tools/tmp/libonewarp.so tools/tmp/libmw4.so dynenv_b200/libdynenv_b200.so tools/tmp/libmw16.so
for lib in dynenv_b200/libdynenv_b200.so tools/tmp/libmw16.so; do
echo == $lib
DYNENV_B200_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv -k regex:rc_phys_slow -s 600 -c 3 --log-file gpurun_out/s5.csv python tools/gpu_rc_variant.py 122 > /dev/null 2>&1
grep -v "^==" gpurun_out/s5.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>5 and r[0]!='ID': print(r[0], r[-3][:60], r[-1])
"
done 2>&1 | tee gpurun_out/s5.log
