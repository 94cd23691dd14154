#!/bin/bash
( timeout 900 python -m pytest tests/test_gpu_robocup.py tests/test_gpu_api.py tests/test_gpu_determinism.py -m gpu -x -q 2>&1 | tail -3 )
export SPARSE=1
N=130 ./gpu_var.sh dynenv_b200/libdynenv_b200.so
