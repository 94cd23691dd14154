#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_robocup.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2c_pytest.log 2>&1
tail -5 gpurun_out/r2c_pytest.log
python tools/gpu_rc_dbg.py 121 2>&1 | tail -6
N=${N:-100} bash tools/gpu/gpu_var.sh dynenv_b200/libdynenv_b200.so
