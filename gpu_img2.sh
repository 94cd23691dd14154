#!/bin/bash
python -m pytest tests/test_gpu_image.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -2
DYNENV_B200_LIB=$PWD/dynenv_b200/libdynenv_b200.so python tools/gpu_raster_time.py 23 2>&1 | tail -1
