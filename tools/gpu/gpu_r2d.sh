#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_robocup.py tests/test_gpu_api.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2d_pytest.log 2>&1
tail -5 gpurun_out/r2d_pytest.log
N=${N:-240} bash tools/gpu/gpu_var.sh dynenv_b200/libdynenv_b200.so
