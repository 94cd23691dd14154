#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_robocup.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2b_pytest.log 2>&1
tail -15 gpurun_out/r2b_pytest.log
N=100 bash tools/gpu/gpu_var.sh dynenv_b200/libdynenv_ref.so dynenv_b200/libdynenv_b200.so
