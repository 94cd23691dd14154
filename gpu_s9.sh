#!/bin/bash
for lib in dynenv_b200/libdynenv_b200.so tools/tmp/libnobig.so; do
  DYNENV_B200_LIB=$PWD/$lib python tools/gpu_raster_time.py 23 2>&1 | tail -1
done
