#!/bin/bash
# GPU box: refresh of the RoboCup part of profiles/r02_* after the last kernel changes (see gpu_prof_r02.sh for the whole set)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_robocup.json 2> gpurun_out/bench_reference_robocup.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_robocup.json 2> gpurun_out/bench_robocup.err; tail -c 300 gpurun_out/bench_robocup.json
python bench.py --steps 240 --warmup 0 --fast-forward 0 --no-cpu --no-other > gpurun_out/bench_robocup_episode.json 2> gpurun_out/bench_robocup_episode.err
python bench.py --steps 20 --warmup 5 --dense-obs --no-cpu --no-other > gpurun_out/bench_robocup_dense.json 2> gpurun_out/bench_robocup_dense.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-other > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_robocup.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-other > gpurun_out/ncu_l1.log 2>&1
export SPARSE=1
python tools/gpu_rc_variant.py 122 > gpurun_out/plain.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:rc_phys_kernel -s 602 -c 1 -o gpurun_out/prof_rc_phys -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_phys_slow -s 601 -c 1 -o gpurun_out/prof_rc_phys_slow -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obsp_kernel -s 2408 -c 1 -o gpurun_out/prof_rc_obsl -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_obsp_kernel -s 2409 -c 1 -o gpurun_out/prof_rc_obsp -f python tools/gpu_rc_variant.py 122 > gpurun_out/ncu_f4.log 2>&1
}
tail -1 gpurun_out/plain.log
ls -la gpurun_out/prof_rc_phys*.ncu-rep gpurun_out/prof_rc_obs[lp].ncu-rep
