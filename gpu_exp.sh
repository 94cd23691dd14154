#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
