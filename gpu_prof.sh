#!/bin/bash
# GPU box: tests, bench, then ncu (launch list + full capture of the step kernel)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --cpu-steps 300 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu1.log 2>&1
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_step -s 3 -c 1 -o gpurun_out/prof -f python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
