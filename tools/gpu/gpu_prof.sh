#!/bin/bash
# GPU box: tests, smoke, the three benches + reference arms, launch lists and ncu --set full captures of every step kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
for w in robocup driving image; do
  python bench.py --impl reference --workload $w --warmup 1 > gpurun_out/bench_reference_$w.json 2> gpurun_out/bench_reference_$w.err; tail -c 300 gpurun_out/bench_reference_$w.json
  python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 600 gpurun_out/bench_$w.json
done
# launch lists (gpu__time_duration per launch) of short runs
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_robocup.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_l1.log 2>&1
python bench.py --workload driving --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_driving.csv python bench.py --workload driving --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_l2.log 2>&1
python bench.py --workload image --envs 2048 --steps 2 --warmup 2 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_image.csv python bench.py --workload image --envs 2048 --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_l3.log 2>&1
# full captures: RoboCup and Driving kernels at step 100 of the episode (contacts present), raster kernel
python bench.py --steps 2 --warmup 100 --no-cpu > gpurun_out/plain.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:rc_phys -s 502 -c 1 -o gpurun_out/prof_rc_phys -f python bench.py --steps 2 --warmup 100 --no-cpu > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obs -s 502 -c 1 -o gpurun_out/prof_rc_obs -f python bench.py --steps 2 --warmup 100 --no-cpu > gpurun_out/ncu_f2.log 2>&1
}
python bench.py --workload driving --steps 2 --warmup 100 --no-cpu > gpurun_out/plain.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:drv_phys -s 100 -c 1 -o gpurun_out/prof_drv_phys -f python bench.py --workload driving --steps 2 --warmup 100 --no-cpu > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:drv_obs -s 100 -c 1 -o gpurun_out/prof_drv_obs -f python bench.py --workload driving --steps 2 --warmup 100 --no-cpu > gpurun_out/ncu_f4.log 2>&1
}
python bench.py --workload image --envs 1024 --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_raster -s 4 -c 1 -o gpurun_out/prof_rc_raster -f python bench.py --workload image --envs 1024 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_f5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
