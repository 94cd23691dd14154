#!/bin/bash
# round 2: flattened observation kernel -- parity + timing
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_robocup.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2a_pytest.log 2>&1
cat gpurun_out/r2a_pytest.log
timeout 300 python tools/gpu_rc_variant.py 100 > gpurun_out/r2a_variant.log 2>&1; cat gpurun_out/r2a_variant.log
