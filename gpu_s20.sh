#!/bin/bash
export SPARSE=1
ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio --clock-control none --csv -k regex:rc_phys_slow -s 600 -c 5 --log-file gpurun_out/s20.csv python tools/gpu_rc_variant.py 122 > /dev/null 2>&1
grep -v "^==" gpurun_out/s20.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>5 and r[0]!='ID': print(r[0], r[-3][:70], r[-1])
"
