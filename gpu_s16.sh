#!/bin/bash
python -m pytest tests/test_gpu_api.py -m gpu -q -x -k "arranger or sparse" 2>&1 | tail -3
python tools/gpu_arr_time.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:arr_ python tools/gpu_arr_time.py 2>/dev/null | grep arr_ | awk -F'","' '{print $5, $NF}' | tail -8
