#!/bin/bash
# GPU box: tests, bench, then ncu (launch list + full capture of the two step kernels)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_phys -s 60 -c 1 -o gpurun_out/prof_phys -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obs -s 60 -c 1 -o gpurun_out/prof_obs -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
