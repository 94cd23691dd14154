#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_image.py "tests/test_gpu_robocup.py::test_noise_statistics_every_list_and_false_positives" -m gpu -x -q -s 2>&1 | tail -25 ) > gpurun_out/s13_pytest.log 2>&1; cat gpurun_out/s13_pytest.log
