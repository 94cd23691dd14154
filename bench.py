#!/usr/bin/env python
"""Benchmark of the rollout hot path: env-steps/s of RoboCup 5v5 (50 substeps + 5 observation ticks x 10
observers per env-step), PARTIAL observation, REALISTIC noise magnitude 2 (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl reference] [--workload robocup|driving]

`--workload driving` measures BASELINE.json configs[2] instead (DrivingEnvironment, 10 cars, PARTIAL, continuous
actions, REALISTIC noise 2: 10 substeps + 1 observation x 10 observers per env-step); the default is configs[1].

One process per GPU (torchrun for N > 1); environments are sharded by contiguous env-id range with no
data-path collective (weak scaling: E environments per GPU).  Prints ONE JSON line on rank 0.

`--impl reference` times the CPU implementation of the same path on the host cores: the reference is
pure Python over pymunk, neither of which exists on the GPU box, so this is the repo's C restatement of
it (oracle/rc_oracle.c, pinned against the reference's own Python by tests/golden), one environment per
host thread.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "env-steps/s"
FLOP_PER_ENV_STEP = 1.0e6  # SURVEY.md 8d contract figure (both workloads)
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


class RoboCupWorkload(object):
    """BASELINE.json configs[1]"""
    name = "robocup"
    metric = "env-steps/sec (RoboCup 5v5, 50 substeps)"
    workload = "RoboCup 5v5 PARTIAL REALISTIC noise 2, canFall, auto-reset"
    P, act = 10, 4
    launches_per_step = {"phys": 5, "obs": 5, "finish": 1}
    prefix = "rc"
    cpu_steps, ref_steps, cpu_warm = 20000, 1000, 50  # env-steps per host thread: cpu_baseline sample, reference-arm step, warm-up
    default_steps = 240  # one whole episode (12000 / 50): contacts, falls and penalties build up over an episode

    def make_sim(self, E, local, rank, full=False):
        import dynenv_b200
        obs_type = dynenv_b200.ObservationType.FULL if full else dynenv_b200.ObservationType.PARTIAL
        return dynenv_b200.BatchedRoboCup(E, 5, obs_type, dynenv_b200.NoiseType.REALISTIC, 2.0, canFall=True,
                                          useObsRewards=True, auto_reset=True, seed=42, device=local, env_id_offset=rank * E)

    def make_actions(self, E, g, dev):
        """int uniform move [0,5) turn [0,3) kick [0,3) head [0,7)"""
        import torch
        hi = torch.tensor([5, 3, 3, 7], device=dev)
        return (torch.rand((E, self.P, 4), generator=g, device=dev) * hi).floor().contiguous()

    def config(self, L):
        return {"substeps": L.RC_SUBSTEPS, "obs_ticks": L.RC_NTICK, "observers": self.P}

    def algorithmic_bytes_per_env_step(self):
        """step-boundary HBM traffic of one env-step: state record in + out, actions in,
        5 x 10 observation records, result record."""
        from dynenv_b200 import layout as L
        return 4 * (2 * L.RC_STATE_WORDS + self.P * 4 + L.RC_NTICK * self.P * L.RC_OBS_WORDS + L.RC_INFO_WORDS)

    def kernel_bytes_per_env(self):
        """algorithmic HBM bytes per environment of ONE launch of each kernel class"""
        from dynenv_b200 import layout as L
        rec = 4 * L.RC_STATE_WORDS
        return {"phys": 2 * rec + 4 * self.P * 4,                                   # record in + out, actions
                "obs": rec + 4 * self.P * L.RC_OBS_WORDS + 4 * (L.RC_S_STEP_END - L.RC_S_STEP_LSUM),  # record in, 10 records out
                "finish": 2 * rec + 4 * L.RC_INFO_WORDS}

    def cpu_env(self, i):
        from oracle import oracle as O
        import ctypes as C
        env = O.RcOracle(5, 1, 1, 2.0, False, seed=42, env_id=i)
        ref = C.byref(env._out)
        return env, lambda a: env._lib.rc_step(env._h, a.ctypes.data, ref)

    def cpu_actions(self, rng, n_steps):
        return np.stack([rng.integers(0, 5, (n_steps, self.P)), rng.integers(0, 3, (n_steps, self.P)),
                         rng.integers(0, 3, (n_steps, self.P)), rng.integers(0, 7, (n_steps, self.P))], -1).astype(np.float64)


class DrivingWorkload(object):
    """BASELINE.json configs[2]"""
    name = "driving"
    metric = "env-steps/sec (Driving 10 cars, 10 substeps)"
    workload = "Driving 10 cars PARTIAL continuous actions REALISTIC noise 2, auto-reset"
    P, act = 10, 2
    launches_per_step = {"phys": 1, "obs": 1, "finish": 0}
    prefix = "drv"
    cpu_steps, ref_steps, cpu_warm = 10000, 1000, 50
    default_steps = 600  # one whole episode (6000 / 10)

    def make_sim(self, E, local, rank, full=False):
        import dynenv_b200
        obs_type = dynenv_b200.ObservationType.FULL if full else dynenv_b200.ObservationType.PARTIAL
        return dynenv_b200.BatchedDriving(E, 10, obs_type, dynenv_b200.NoiseType.REALISTIC, 2.0, continuousActions=True,
                                          auto_reset=True, seed=42, device=local, env_id_offset=rank * E)

    def make_actions(self, E, g, dev):
        """float uniform [-3, 3) acceleration and steering"""
        import torch
        return (torch.rand((E, self.P, 2), generator=g, device=dev) * 6.0 - 3.0).contiguous()

    def config(self, L):
        return {"substeps": L.DRV_SUBSTEPS, "obs_ticks": 1, "observers": self.P}

    def algorithmic_bytes_per_env_step(self):
        from dynenv_b200 import layout as L
        return 4 * (2 * L.DRV_STATE_WORDS + self.P * 2 + self.P * L.DRV_OBS_WORDS + L.DRV_INFO_WORDS)

    def kernel_bytes_per_env(self):
        from dynenv_b200 import layout as L
        rec = 4 * L.DRV_STATE_WORDS
        return {"phys": 2 * rec + 4 * self.P * 2,
                "obs": 2 * rec + 4 * self.P * L.DRV_OBS_WORDS + 4 * L.DRV_INFO_WORDS,  # record in + out, 10 records, result record
                "finish": 0}

    def cpu_env(self, i):
        from oracle import oracle as O
        import ctypes as C
        env = O.DrvOracle(10, 1, 1, 2.0, True, seed=42, env_id=i)
        ref = C.byref(env._out)
        return env, lambda a: env._lib.drv_step(env._h, a.ctypes.data, ref)

    def cpu_actions(self, rng, n_steps):
        return rng.uniform(-3, 3, (n_steps, self.P, 2))


class _ImageSim(object):
    """RoboCup simulator in IMAGE mode + the raster kernel over every record of the step, in environment chunks that
    fit one reusable plane buffer (15.36 MB of bit-planes per environment-step)"""

    def __init__(self, sim, chunk):
        import torch
        from dynenv_b200 import layout as L
        self.sim, self.chunk = sim, chunk
        self.out = sim.out
        self.buf = torch.empty((chunk, L.RC_NTICK, sim.P, L.RC_IMG_PLANES, L.RC_IMG_H, L.RC_IMG_WORDS_PER_ROW), dtype=torch.int32,
                               device=sim.device)
        self.timing = False
        self.ev = None

    def reset(self):
        return self.sim.reset()

    def step(self, actions):
        import torch
        out = self.sim.step(actions)
        if self.timing:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        for b in range(0, self.sim.nEnvs, self.chunk):
            e = min(b + self.chunk, self.sim.nEnvs)
            self.sim.render_images(b, e, out=self.buf[:e - b])
        if self.timing:
            self.ev[1].record()
        return out

    def debug_timing(self, enable=True):
        self.timing = bool(enable)
        self.sim.debug_timing(enable)

    def last_kernel_ms(self):
        k = self.sim.last_kernel_ms()
        n = (self.sim.nEnvs + self.chunk - 1) // self.chunk
        self.ev[1].synchronize()
        k["raster"] = (self.ev[0].elapsed_time(self.ev[1]), n)
        return k

    def last_launch_count(self):
        return self.sim.last_launch_count() + (self.sim.nEnvs + self.chunk - 1) // self.chunk


class ImageWorkload(RoboCupWorkload):
    """BASELINE.json configs[3]"""
    name = "image"
    metric = "env-steps/sec (RoboCup 5v5 IMAGE, 50 substeps + 100 camera images)"
    workload = "RoboCup 5v5 IMAGE (top / bottom camera raster, 8 x 480 x 640 bit-planes per robot and tick) allowHeadTurn REALISTIC noise 2, auto-reset"
    launches_per_step = {"phys": 5, "obs": 5, "finish": 1, "raster": 1}
    chunk = 1024
    default_envs = 16384
    cpu_steps, ref_steps, cpu_warm = 16, 4, 1  # an env-step draws 100 images with cv2 (~0.3 s)
    default_steps = 10

    def make_sim(self, E, local, rank, full=False):
        import dynenv_b200
        sim = dynenv_b200.BatchedRoboCup(E, 5, dynenv_b200.ObservationType.IMAGE, dynenv_b200.NoiseType.REALISTIC, 2.0, allowHeadTurn=True,
                                         canFall=True, useObsRewards=True, auto_reset=True, seed=42, device=local, env_id_offset=rank * E)
        self.launches_per_step = dict(self.launches_per_step, raster=(E + self.chunk - 1) // self.chunk)
        self.E = E
        return _ImageSim(sim, min(self.chunk, E))

    def make_actions(self, E, g, dev):
        """three categorical columns as floats + head turn uniform [-3, 3)"""
        import torch
        a = torch.rand((E, self.P, 4), generator=g, device=dev)
        a[..., :3] = (a[..., :3] * torch.tensor([5, 3, 3], device=dev)).floor()
        a[..., 3] = a[..., 3] * 6.0 - 3.0
        return a.contiguous()

    def config(self, L):
        return dict(RoboCupWorkload.config(self, L), images_per_env_step=2 * L.RC_NTICK * self.P, render_chunk_envs=self.chunk)

    def algorithmic_bytes_per_env_step(self):
        from dynenv_b200 import layout as L
        return RoboCupWorkload.algorithmic_bytes_per_env_step(self) + 4 * L.RC_NTICK * self.P * (2 * L.RC_RAW_WORDS + L.RC_IMG_WORDS)

    def kernel_bytes_per_env(self):
        from dynenv_b200 import layout as L
        kb = RoboCupWorkload.kernel_bytes_per_env(self)
        kb["obs"] += 4 * self.P * L.RC_RAW_WORDS
        # one raster launch covers `chunk` environments: raw records in, bit-planes out (bytes per environment of the chunk
        # x chunk / E so that kb * E is the launch's traffic)
        kb["raster"] = 4 * L.RC_NTICK * self.P * (L.RC_RAW_WORDS + L.RC_IMG_WORDS) * min(self.chunk, self.E) / self.E
        return kb

    def cpu_env(self, i):
        """the C oracle's step with the raw detection lists captured, then the reference's drawing calls (real cv2) for
        every robot and tick"""
        import ctypes as C
        from oracle import oracle as O, raster_oracle
        env = O.RcOracle(5, 2, 1, 2.0, True, seed=42, env_id=i)
        env.capture_raw()
        ref = C.byref(env._out)

        def step(a):
            env._lib.rc_step(env._h, a.ctypes.data, ref)
            for t in range(5):
                for ag in range(self.P):
                    raster_oracle.render(env.raw_vision(t, ag))
        return env, step

    def cpu_actions(self, rng, n_steps):
        a = RoboCupWorkload.cpu_actions(self, rng, n_steps)
        a[..., 3] = rng.uniform(-3, 3, a.shape[:-1])
        return a


WORKLOADS = {"robocup": RoboCupWorkload, "driving": DrivingWorkload, "image": ImageWorkload}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)"""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_env_steps_per_s(wl, n_steps, threads=None):
    """the C oracle, one environment per host thread (ctypes releases the GIL), all threads stepping
    n_steps env-steps; returns (env-steps/s over all threads, threads)"""
    import ctypes as C
    from oracle import oracle as O
    O.lib()
    threads = threads or os.cpu_count() or 1
    out = [0.0] * threads
    barrier = threading.Barrier(threads)

    def work(i):
        env, step_fn = wl.cpu_env(i)
        env.reset()
        rng = np.random.default_rng(1234 + i)
        acts = np.ascontiguousarray(wl.cpu_actions(rng, n_steps), dtype=np.float64)
        barrier.wait()
        t0 = time.perf_counter()
        for s in range(n_steps):
            step_fn(acts[s])
            if env._out.done:  # the vec-env's auto-reset (utils/subproc_vec_env.py:17-23)
                env.reset()
        out[i] = time.perf_counter() - t0

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return threads * n_steps / max(out), threads, max(out)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]()
    METRIC, WORKLOAD = wl.metric, wl.workload
    # each "step" = every host thread advancing a bounded sample of env-steps of the same workload
    per_step = wl.ref_steps
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_env_steps_per_s(wl, wl.cpu_warm)
    t_total, n_total, threads = 0.0, 0, 0
    for _ in range(args.steps):
        rate, threads, dt = cpu_env_steps_per_s(wl, per_step)
        t_total += dt
        n_total += threads * per_step
    value = n_total / t_total
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "envs": threads, "env_steps_per_step": threads * per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d host threads x %d env-steps per bench step, C restatement of the reference "
                                       "(pymunk and the reference's Python are not installable on this box)" % (threads, per_step)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import dynenv_b200
    from dynenv_b200 import layout as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]()
    E = args.envs or getattr(wl, "default_envs", 65536)
    METRIC, WORKLOAD, P = wl.metric, wl.workload, wl.P
    sim = wl.make_sim(E, local, rank, full=args.debug_full_obs)
    sim.reset()
    # synthetic actions, generator seed 1234
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    # fresh uniform actions every step: a cycle biases the play (every robot repeats its walk, they drift together and
    # collide: a 32-step cycle costs 12 % env-steps/s, tools/gpu_e2e_parts.py), so the pool covers the whole run
    n_pool = min(args.steps + args.warmup, 640)
    pool = [wl.make_actions(E, g, dev) for _ in range(n_pool)]
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for w in range(args.warmup):
        sim.step(pool[w % n_pool])
    pool = pool[args.warmup % n_pool:] + pool[:args.warmup % n_pool]
    barrier()
    sim.debug_timing(True)
    kms = {k: [0.0, 0] for k in wl.launches_per_step}
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    t_start.record(stream)
    for s in range(args.steps):
        evs[s][0].record(stream)
        sim.step(pool[s % n_pool])
        evs[s][1].record(stream)
        launches += sim.last_launch_count()
        if s % 4 == 3 or s == args.steps - 1:  # per-kernel durations (waits for that step; every 4th only)
            for k, (ms, n) in sim.last_kernel_ms().items():
                kms[k][0] += ms
                kms[k][1] += n
    t_end.record(stream)
    barrier()
    sim.debug_timing(False)
    clocks = sampler.stop()
    ms_total = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    diag = int(sim.out.diag.max().item())

    # end to end through the host-facing call: pinned host actions in, rewards + done out, every step
    # the timed steps' own actions, from pinned host memory (at most ~3.5 GB of it per rank; cycled beyond that)
    n_host = min(n_pool, args.steps, max(32, int(3.5e9 // (E * P * wl.act * 4))))
    h_act = [p.cpu().pin_memory() for p in pool[:n_host]]
    n_pool = len(h_act)
    d_act = torch.empty((E, P, wl.act), device=dev)
    # rewards [maxAgents] and done are the first words of every environment's result record: one strided copy
    # (two copies of the `rewards` / `done` views cost 0.45-1.0 ms per step, tools/gpu_e2e_parts.py)
    n_res = (L.DRV_I_DONE if args.workload == "driving" else L.RC_I_DONE) + 1
    h_res = torch.empty((E, n_res), dtype=torch.float32).pin_memory()
    # the same window of the episode as the device-resident measurement: reset, the same warm-up steps, then the timed
    # steps on their own actions (step times grow over an episode; a later window would not be comparable)
    sim.reset()
    for w in range(args.warmup):
        sim.step(pool[(len(pool) - (args.warmup % len(pool)) + w) % len(pool)])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(args.steps):
        d_act.copy_(h_act[s % n_pool], non_blocking=True)
        out = sim.step(d_act)
        h_res.copy_(out.info[:, :n_res], non_blocking=True)
        stream.synchronize()
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, kernel_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, kernel_ms = [float(x) for x in t.tolist()]
    if rank == 0:
        peaks, which = measured_peaks()
        import ctypes
        from dynenv_b200 import _lib as _l
        tf = ctypes.c_float(0)
        fp32_peak = float(tf.value) if _l.lib().dynenv_debug_fp32_peak(local, ctypes.byref(tf)) == 0 else None
        total_envs = E * world
        value = total_envs * args.steps / (ms_total * 1e-3)
        e2e = total_envs * args.steps / (ms_e2e * 1e-3)
        bytes_launch = wl.algorithmic_bytes_per_env_step() * E
        kb = wl.kernel_bytes_per_env()
        per_launch = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in kms.items()}  # ms per launch
        per_step = {k: per_launch[k] * wl.launches_per_step[k] for k in per_launch}
        dom = max(per_step, key=per_step.get)  # dominant kernel class of the step
        achieved = kb[dom] * E / (per_launch[dom] * 1e-3) / 1e9 if per_launch[dom] else 0.0
        step_gbs = bytes_launch / (kernel_ms * 1e-3) / 1e9
        traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get("%s_%s_kernel" % (wl.prefix, dom))
            if tr:  # scaled to the environments one launch of this run covers
                traffic = tr["dram_bytes_per_launch"] * (min(getattr(wl, "chunk", E), E) if dom == "raster" else E) / tr["envs"]
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict({"workload": WORKLOAD, "envs_per_gpu": E, "envs_total": total_envs,
                                "episode_steps": [args.warmup, args.warmup + args.steps],  # window of the timed steps in the episode(s)
                                "l2": "state + outputs per step (%.1f GB) exceed the 126 MB L2" % (bytes_launch / 1e9)},
                               **wl.config(L)),
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": E * P * wl.act * 4,
                        "d2h_bytes_per_step": E * n_res * 4,
                        "note": "observations stay device-resident for the learner; rewards + done return to the host"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": which,
                             "kernel": "%s_%s_kernel" % (wl.prefix, dom), "kernel_ms_per_launch": per_launch[dom],
                             "algorithmic_bytes_per_launch": kb[dom] * E,
                             "launch_ms": per_launch, "share_of_step": {k: v / max(sum(per_step.values()), 1e-9) for k, v in per_step.items()},
                             "step": {"ms": kernel_ms, "algorithmic_bytes_per_env_step": wl.algorithmic_bytes_per_env_step(),
                                      "achieved_gbs": step_gbs, "frac": step_gbs / peaks["hbm_gbs"]},
                             "note": "the step is instruction-issue bound (FP32/INT pipes), not HBM bound; see DESIGN.md section 4",
                             "fp32": {"flop_per_env_step": FLOP_PER_ENV_STEP,
                                      "achieved_tflops": FLOP_PER_ENV_STEP * E / (kernel_ms * 1e-3) / 1e12,
                                      "nominal_peak_tflops": FP32_NOMINAL_TFLOPS,
                                      "measured_peak_tflops": fp32_peak,  # FMA probe of the library, this run
                                      "frac": FLOP_PER_ENV_STEP * E / (kernel_ms * 1e-3) / 1e12 / (fp32_peak or FP32_NOMINAL_TFLOPS)}},
                "diag_flags": diag}
        if not args.no_cpu and world == 1:
            cpu_steps = args.cpu_steps or wl.cpu_steps
            rate, threads, dt = cpu_env_steps_per_s(wl, cpu_steps)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d host threads x %d env-steps of the same workload (%.1f s), C restatement "
                                              "of the reference" % (threads, cpu_steps, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: one episode of the workload; reference arm 3)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs", type=int, default=None, help="environments per GPU (default 65536; image workload 16384)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="robocup", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--debug-full-obs", action="store_true", help="(profiling aid) FULL observations: physics and rules only")
    ap.add_argument("--cpu-steps", type=int, default=None, help="env-steps per host thread of the cpu_baseline sample")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if args.impl == "reference" else WORKLOADS[args.workload].default_steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
