#!/bin/bash
# Shared-memory bounds check of every kernel (compute-sanitizer is not available on the GPU pool):
#   here:        tools/gpu_bounds.sh build     -> tools/tmp/libbounds.so (-DDYNENV_BOUNDS_CHECK)
#   on the box:  tools/gpu_bounds.sh           -> the GPU test-suite on that library; tests/conftest.py prints the count
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  mkdir -p tools/tmp
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -shared -Xcompiler -fPIC \
       -DDYNENV_BOUNDS_CHECK -o tools/tmp/libbounds.so dynenv_b200/csrc/dynenv_api.cu dynenv_b200/csrc/rc_slow.cu
  exit $?
fi
DYNENV_B200_LIB=$PWD/tools/tmp/libbounds.so python -m pytest tests -m gpu -q -s -k "not host_api" 2>&1 | tail -8
