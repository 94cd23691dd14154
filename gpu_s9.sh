#!/bin/bash
for lib in tools/tmp/libr512.so tools/tmp/libr384.so tools/tmp/libr128.so; do
  DYNENV_B200_LIB=$PWD/$lib python tools/gpu_raster_time.py 23 2>&1 | tail -1
done
