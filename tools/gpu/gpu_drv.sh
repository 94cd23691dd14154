#!/bin/bash
# GPU box: Driving parity tests with full mismatch output
mkdir -p gpurun_out
python -m pytest tests/test_gpu_driving.py -m gpu -q 2>&1 | tail -150 > gpurun_out/pytest_drv.log
tail -5 gpurun_out/pytest_drv.log
