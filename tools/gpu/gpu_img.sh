#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_image.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_img.log; tail -12 gpurun_out/pytest_img.log
