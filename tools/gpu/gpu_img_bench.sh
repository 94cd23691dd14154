#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload image --steps 6 --warmup 3 > gpurun_out/bench_img.json 2> gpurun_out/bench_img.err; tail -c 3000 gpurun_out/bench_img.json; tail -3 gpurun_out/bench_img.err
python bench.py --workload image --envs 2048 --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_img.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_raster -s 4 -c 1 -o gpurun_out/prof_raster -f python bench.py --workload image --envs 2048 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_img.log 2>&1
tail -2 gpurun_out/ncu_img.log
