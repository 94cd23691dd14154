#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_driving.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/s15_pytest.log 2>&1; cat gpurun_out/s15_pytest.log
for lib in tools/tmp/libdrv_onewarp.so tools/tmp/libdrv_w4.so dynenv_b200/libdynenv_b200.so tools/tmp/libdrv_w16.so tools/tmp/libdrv_w32.so; do
  DYNENV_B200_LIB=$PWD/$lib timeout 300 python tools/gpu_drv_variant.py 330 2>&1 | tail -1
done | tee gpurun_out/s15_var.log
