#!/bin/bash
# GPU box: everything profiles/r02_* is made from -- tests, smoke, the benches + reference arms, launch lists (ncu
# gpu__time_duration per launch) and one ncu --set full capture of every step kernel mid-episode (tools/make_profiles.py r02)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_robocup.json 2> gpurun_out/bench_reference_robocup.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_robocup.json 2> gpurun_out/bench_robocup.err; tail -c 400 gpurun_out/bench_robocup.json
python bench.py --steps 240 --warmup 0 --fast-forward 0 --no-cpu --no-other > gpurun_out/bench_robocup_episode.json 2> gpurun_out/bench_robocup_episode.err
python bench.py --steps 20 --warmup 5 --dense-obs --no-cpu --no-other > gpurun_out/bench_robocup_dense.json 2> gpurun_out/bench_robocup_dense.err
for w in driving image; do
  python bench.py --impl reference --workload $w --warmup 1 > gpurun_out/bench_reference_$w.json 2> gpurun_out/bench_reference_$w.err
  python bench.py --workload $w --no-other > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 300 gpurun_out/bench_$w.json
done
# launch lists of short runs
python bench.py --steps 3 --warmup 3 --no-cpu --no-other > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_robocup.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-other > gpurun_out/ncu_l1.log 2>&1
python bench.py --workload driving --steps 3 --warmup 3 --no-cpu --no-other > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_driving.csv python bench.py --workload driving --steps 3 --warmup 3 --no-cpu --no-other > gpurun_out/ncu_l2.log 2>&1
python bench.py --workload image --envs 2048 --steps 2 --warmup 2 --no-cpu --no-other > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_image.csv python bench.py --workload image --envs 2048 --steps 2 --warmup 2 --no-cpu --no-other > gpurun_out/ncu_l3.log 2>&1
# full captures at step 120 of the RoboCup episode (sparse observation stream, as in the bench): launch indices per step are
# 5 x rc_phys_kernel, 5 x rc_phys_slow_kernel, 20 x rc_obsp_kernel (lines, points on the caller's stream, then the handed-over ones')
export SPARSE=1
python tools/gpu_rc_variant.py 122 > gpurun_out/plain.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:rc_phys_kernel -s 602 -c 1 -o gpurun_out/prof_rc_phys -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_phys_slow -s 601 -c 1 -o gpurun_out/prof_rc_phys_slow -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obsp_kernel -s 2408 -c 1 -o gpurun_out/prof_rc_obsl -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obsp_kernel -s 2409 -c 1 -o gpurun_out/prof_rc_obsp -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_f4.log 2>&1
}
python bench.py --workload driving --steps 2 --warmup 100 --fast-forward 0 --no-cpu --no-other > gpurun_out/plain.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:drv_phys -s 100 -c 1 -o gpurun_out/prof_drv_phys -f python bench.py --workload driving --steps 2 --warmup 100 --fast-forward 0 --no-cpu --no-other > gpurun_out/ncu_f5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:drv_obs -s 100 -c 1 -o gpurun_out/prof_drv_obs -f python bench.py --workload driving --steps 2 --warmup 100 --fast-forward 0 --no-cpu --no-other > gpurun_out/ncu_f6.log 2>&1
}
python tools/gpu_raster_time.py 23 > gpurun_out/raster_time.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rc_raster -s 2 -c 1 -o gpurun_out/prof_rc_raster -f python tools/gpu_raster_time.py 23 > gpurun_out/ncu_f7.log 2>&1
cat gpurun_out/raster_time.log | tail -1
ls -la gpurun_out/prof_rc_*.ncu-rep gpurun_out/prof_drv_*.ncu-rep | tail -8
