#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_robocup.py tests/test_gpu_api.py -m gpu -x -q -s 2>&1 | tail -15 ) > gpurun_out/s6_pytest.log 2>&1
tail -6 gpurun_out/s6_pytest.log
N=130 ./gpu_var.sh dynenv_b200/libdynenv_b200.so
cp gpurun_out/var.log gpurun_out/s6_var_dense.log
SPARSE=1 N=130 ./gpu_var.sh dynenv_b200/libdynenv_b200.so
