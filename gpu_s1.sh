#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_image.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s1_pytest_img.log; cat gpurun_out/s1_pytest_img.log
for lib in tools/tmp/libraster_static.so dynenv_b200/libdynenv_b200.so; do
  echo == $lib; DYNENV_B200_LIB=$PWD/$lib python tools/gpu_raster_prof.py 2>&1 | head -1
done | tee gpurun_out/s1_raster.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-other > gpurun_out/s1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s1_launches_robocup.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-other > gpurun_out/s1_ncu_l1.log 2>&1
tail -1 gpurun_out/s1_plain.log | cut -c1-400
