#!/bin/bash
mkdir -p gpurun_out
N=130 ./gpu_var.sh tools/tmp/libslow4.so tools/tmp/libslow8.so tools/tmp/libslow16.so dynenv_b200/libdynenv_b200.so
for lib in tools/tmp/libslow8.so; do
DYNENV_B200_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv -k regex:rc_phys_slow -s 600 -c 5 --log-file gpurun_out/s3_slow8.csv python tools/gpu_rc_variant.py 122 > /dev/null 2>&1
done
grep -v "^==" gpurun_out/s3_slow8.csv | cut -d, -f5,13- | tail -12
