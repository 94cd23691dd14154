"""Dev tool (here): turns what tools/gpu/gpu_prof*.sh brought back in gpurun_out/ into the tracked summaries under profiles/:
  profiles/rNN_ncu_<kernel>.txt       selected metrics of the ncu --set full capture + per-function instruction / stall table
  profiles/rNN_launches_<workload>.txt per-kernel launch counts, mean duration, share of the step kernels
  profiles/rNN_bench_<workload>.json   the bench lines (GPU arm and reference arm)
  profiles/rNN_traffic.json            DRAM bytes per launch of every captured kernel (bench.py reads it for roofline.traffic)
usage: python tools/make_profiles.py r01"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT, SRC = "profiles", "gpurun_out"
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_blocks", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
KERNELS = {"rc_phys": ("rc_phys_kernel", "65536 RoboCup 5v5 environments, 10 substeps each; step 120 of the episode", 65536 * 10),
           "rc_phys_slow": ("rc_phys_slow", "the environments rc_phys_kernel handed over in one block of 10 substeps (6-9 % of 65536), 16 per thread block; step 120", 65536 * 10 * 0.08),
           "rc_obsl": ("rc_obsp_kernelILb1", "65536 RoboCup 5v5 environments, the field lines of one observation tick (10 observers each); step 120, sparse stream", 65536 * 10),
           "rc_obsp": ("rc_obsp_kernelILb0", "65536 RoboCup 5v5 environments, the point objects of one observation tick (10 observers each); step 120, sparse stream", 65536 * 10),
           "rc_obs": ("rc_obs_kernel", "65536 RoboCup 5v5 environments, one observation tick (10 observers each); step 100", 65536 * 10),
           "drv_phys": ("drv_phys_kernel", "65536 Driving environments (10 cars), 10 substeps each; step 100 of the episode", 65536 * 10),
           "drv_obs": ("drv_obs_kernel", "65536 Driving environments, observation of 10 cars + rewards / info; step 100", 65536 * 10),
           "rc_raster": ("rc_raster_kernel", "256 RoboCup environments x 5 ticks x 10 robots x 8 planes = 102400 blocks (one plane each), the bench's IMAGE workload at step 23", 102400)}
traffic = {}
BYTE_UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for short, (kern, what, units) in KERNELS.items():
    rep = os.path.join(SRC, "prof_%s.ncu-rep" % short)
    if not os.path.exists(rep):
        continue
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    lines = ["ncu --set full --clock-control none --import-source on, one launch of %s" % kern, "(%s).  Times under ncu are serialised, cold-cache replays." % what, ""]
    dram = 0.0
    for m in METRICS:
        if m in h:
            i = h.index(m)
            lines.append("  %-92s %-16s %s" % (m, u[i], v[i]))
            if m.startswith("dram__bytes"):
                dram += float(v[i]) * BYTE_UNITS.get(u[i], 1)
    traffic[kern] = {"dram_bytes_per_launch": dram, "envs": 256 if short == "rc_raster" else 65536,
                     "source": "profiles/%s_ncu_%s.txt" % (tag, short)}
    env = dict(os.environ, KERNEL=kern, NENV=str(units))
    tab = subprocess.run([sys.executable, "tools/prof_by_func.py", rep], capture_output=True, text=True, env=env).stdout
    lines += ["", "per device function (instructions per unit = per environment-substep / observer-tick / plane; stall samples):", tab]
    open(os.path.join(OUT, "%s_ncu_%s.txt" % (tag, short)), "w").write("\n".join(lines))
    print("wrote", short)
json.dump(traffic, open(os.path.join(OUT, "%s_traffic.json" % tag), "w"), indent=1)

for wl in ("robocup", "driving", "image"):
    p = os.path.join(SRC, "launches_%s.csv" % wl)
    if os.path.exists(p):
        agg = collections.OrderedDict()
        rows = [r for r in csv.reader(open(p)) if len(r) > 10]
        hdr = rows[0]
        ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        for r in rows[1:]:
            try:
                val = float(r[iv].replace(",", ""))
            except ValueError:
                continue
            ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
            a = agg.setdefault(re.sub(r"\(.*", "", r[ik])[:70], [0, 0.0])
            a[0] += 1
            a[1] += ms
        step = sum(v[1] for k, v in agg.items() if re.match(r"(void )?(\w+::)*(rc|drv)_(phys|obs|finish|raster)", k) or "_kernel" in k and ("rc_" in k or "drv_" in k) and "reset" not in k)
        out = ["ncu --metrics gpu__time_duration.sum --clock-control none: bench.py --workload %s (short run: GPU arm + end-to-end steps)" % wl]
        for k, (n, ms) in agg.items():
            mine = ("rc_" in k or "drv_" in k) and "reset" not in k
            out.append("%-72s launches %4d  mean %8.3f ms  sum %9.2f ms  share of step kernels %s" % (k, n, ms / n, ms, ("%.1f%%" % (100 * ms / step)) if mine and step else "-"))
        open(os.path.join(OUT, "%s_launches_%s.txt" % (tag, wl)), "w").write("\n".join(out) + "\n")
        print("wrote launches", wl)
    for arm, name in (("bench_%s.json" % wl, "%s_bench_%s.json" % (tag, wl)), ("bench_reference_%s.json" % wl, "%s_bench_reference_%s.json" % (tag, wl))):
        q = os.path.join(SRC, arm)
        if os.path.exists(q) and os.path.getsize(q):
            open(os.path.join(OUT, name), "w").write(open(q).read().strip().splitlines()[-1] + "\n")
