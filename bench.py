#!/usr/bin/env python
"""Benchmark of the rollout hot path: env-steps/s of RoboCup 5v5 (50 substeps + 5 observation ticks x 10
observers per env-step), PARTIAL observation, REALISTIC noise magnitude 2 (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl reference] [--workload robocup|driving|image]

The timed window starts mid-episode (an untimed fast-forward of 110 RoboCup / 290 Driving steps in front of the W warm-up
steps): contacts, falls and penalties build up over an episode, and the first steps of an episode are its cheapest.
The default RoboCup line also carries short legs of the other BASELINE.json configs (`other_workloads`: Driving, IMAGE,
single-environment latency) and the drop-in API's end-to-end rate (`e2e_dropin`).

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
    launches_per_step = {"phys": 5, "phys_slow": 5, "obs": 5, "finish": 1}  # obs: one tick = lines + point objects, both groups
    kernel_names = {"phys": "rc_phys_kernel", "phys_slow": "rc_phys_slow_kernel", "obs": "rc_obsp_kernel<lines> + rc_obsp_kernel<points>",
                    "finish": "rc_finish_kernel", "raster": "rc_raster_kernel"}
    prefix = "rc"
    cpu_steps, ref_steps, cpu_warm = 20000, 1000, 50  # env-steps per host thread: cpu_baseline sample, reference-arm step, warm-up
    default_steps = 240  # one whole episode (12000 / 50): contacts, falls and penalties build up over an episode
    # DYNENV_FLAG_SPARSE_OBS: counts-prefixed observation stream, rows behind the counts are not written (--dense-obs: zero-filled
    # padded records).  obs_bytes = observation bytes written per env-step, measured on the last timed step (dense: the records).
    sparse_obs = True
    obs_bytes = None
    byte_actions = True  # e2e: the MultiDiscrete action indices travel as uint8 (dynenv_step_u8)

    def make_sim(self, E, local, rank, full=False):
        import dynenv_b200
        obs_type = dynenv_b200.ObservationType.FULL if full else dynenv_b200.ObservationType.PARTIAL
        return dynenv_b200.BatchedRoboCup(E, 5, obs_type, dynenv_b200.NoiseType.REALISTIC, 2.0, canFall=True,
                                          useObsRewards=True, auto_reset=True, seed=42, device=local, env_id_offset=rank * E,
                                          sparse_obs=self.sparse_obs and not full)

    def measure_obs_bytes(self, sim):
        """bytes of observation records one env-step writes: the rows in front of the counts + the 17 count / flag words per
        observer and tick (sparse stream), from the counts of the step just taken; the padded records otherwise"""
        from dynenv_b200 import layout as L
        import torch
        out = getattr(sim, "out", None)
        if not self.sparse_obs or out is None or getattr(out, "counts", None) is None:
            self.obs_bytes = 4 * L.RC_NTICK * self.P * L.RC_OBS_WORDS
            return
        feat = torch.tensor([L.RC_F_BALL, L.RC_F_ROBOT, L.RC_F_POST, L.RC_F_CROSS, L.RC_F_FCROSS, L.RC_F_LINE], device=out.counts.device)
        rows = (out.counts.long() * feat).sum().item() / out.counts.shape[0]
        self.obs_bytes = 4 * (rows + L.RC_NTICK * self.P * (L.RC_NTYPES + 2 + L.RC_MAXP - 1))

    def make_actions(self, E, g, dev):
        """int uniform move [0,5) turn [0,3) kick [0,3) head [0,7)"""
        import torch
        hi = torch.tensor([5, 3, 3, 7], device=dev)
        return (torch.rand((E, self.P, 4), generator=g, device=dev) * hi).floor().contiguous()

    def config(self, L):
        c = {"substeps": L.RC_SUBSTEPS, "obs_ticks": L.RC_NTICK, "observers": self.P}
        if self.name == "robocup":
            c["obs_stream"] = ("counts-prefixed rows (DYNENV_FLAG_SPARSE_OBS): %.1f KB of observations written per env-step, padded "
                               "records would be %.1f KB" % ((self.obs_bytes or 0) / 1e3, 4 * L.RC_NTICK * self.P * L.RC_OBS_WORDS / 1e3)
                               if self.sparse_obs else "padded records, zero-filled behind the counts (--dense-obs)")
        return c

    def _obs_bytes(self):
        from dynenv_b200 import layout as L
        return self.obs_bytes if self.obs_bytes is not None else 4 * L.RC_NTICK * self.P * L.RC_OBS_WORDS

    def algorithmic_bytes_per_env_step(self):
        """step-boundary HBM traffic of one env-step: state record in + out, actions in,
        5 x 10 observation records (sparse stream: the rows and counts actually written), result record."""
        from dynenv_b200 import layout as L
        return 4 * (2 * L.RC_STATE_WORDS + self.P * 4 + L.RC_INFO_WORDS) + self._obs_bytes()

    def kernel_bytes_per_env(self):
        """algorithmic HBM bytes per environment of ONE launch of each kernel class"""
        from dynenv_b200 import layout as L
        rec = 4 * L.RC_STATE_WORDS
        return {"phys": 2 * rec + 4 * self.P * 4,                                   # record in + out, actions
                # one tick: the robot table of the record (21 positions, 20 angles, 10 heads, 10 flags, 7 scalars) twice,
                # 10 records out, the processSeens sums in + out
                "obs": 2 * 4 * 89 + self._obs_bytes() / L.RC_NTICK + 2 * 4 * (L.RC_S_STEP_END - L.RC_S_STEP_LSUM),
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
    kernel_names = {"phys": "drv_phys_kernel", "obs": "drv_obs_kernel"}
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

    def debug_timing(self, enable=True, serialise=False):
        self.timing = bool(enable)
        self.sim.debug_timing(enable, serialise)

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
    launches_per_step = {"phys": 5, "phys_slow": 5, "obs": 5, "finish": 1, "raster": 1}
    chunk = 1024
    sparse_obs = False  # IMAGE mode writes whole records (staged in shared memory)
    byte_actions = False  # allowHeadTurn: the head action is continuous
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
FAST_FORWARD = {"robocup": 110, "driving": 290, "image": 20}  # untimed steps in front of the warm-up: mid-episode


def measure(wl, E, local, rank, world, steps, warmup, ff, e2e=True, per_kernel=True, full=False, sampler=None, arranger=False):
    """One workload on this rank's GPU.  Three passes over the SAME window of the episode (reset, `ff` untimed
    fast-forward steps, `warmup` steps, `steps` timed steps on the same actions):
      1. device-resident: actions already in HBM, CUDA events around the K steps          -> value
      2. end to end: pinned host actions -> H2D -> step -> rewards + done D2H, every step  -> e2e
      3. per kernel class: events bracketing every launch inside the library (untimed)     -> launch_ms
    Returns a dict of raw numbers (this rank's)."""
    import torch
    import torch.distributed as dist
    from dynenv_b200 import layout as L
    dev = torch.device("cuda", local)
    P = wl.P
    sim = wl.make_sim(E, local, rank, full=full)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    # fresh uniform actions for every step of the window (a short cycle biases the play: robots repeat their walk, drift
    # together and collide); the fast-forward steps draw theirs on the fly
    pool = [wl.make_actions(E, g, dev) for _ in range(min(steps, 640))]
    n_pool = len(pool)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def rewind():
        """reset + fast-forward + warm-up: the same episode window for every pass (same seeds, same actions)"""
        sim.reset()
        gf = torch.Generator(device=dev)
        gf.manual_seed(99 + rank)
        for _ in range(ff + warmup):
            sim.step(wl.make_actions(E, gf, dev))
        barrier()

    res = {"E": E, "launches": 0}
    # ---- pass 1: device-resident
    rewind()
    if sampler is not None:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for s in range(steps):
        sim.step(pool[s % n_pool])
        res["launches"] += sim.last_launch_count()
    t1.record(stream)
    barrier()
    if sampler is not None:
        res["clocks"] = sampler.stop()
    res["ms_total"] = t0.elapsed_time(t1)
    res["diag"] = int(sim.out.diag.max().item())
    if hasattr(wl, "measure_obs_bytes"):
        wl.measure_obs_bytes(sim)
    # ---- pass 2: end to end through the host-facing call
    n_res = (L.DRV_I_DONE if wl.name == "driving" else L.RC_I_DONE) + 1
    # MultiDiscrete action spaces (RoboCup here) reach the device as bytes (dynenv_step_u8): what a host-side policy emits
    # are action indices; continuous actions (Driving here, IMAGE's head turn) stay float32
    byte_actions = bool(getattr(wl, "byte_actions", False))
    act_bytes = 1 if byte_actions else 4
    res["h2d"], res["d2h"] = E * P * wl.act * act_bytes, E * n_res * 4
    if e2e:
        n_host = min(n_pool, max(8, int(3.5e9 // (E * P * wl.act * 4))))  # at most ~3.5 GB of pinned actions per rank
        h_act = [(p.to(torch.uint8) if byte_actions else p).cpu().pin_memory() for p in pool[:n_host]]
        d_act = torch.empty((E, P, wl.act), device=dev, dtype=torch.uint8 if byte_actions else torch.float32)
        # rewards [maxAgents] and done are the first words of every environment's result record: one strided copy
        h_res = torch.empty((E, n_res), dtype=torch.float32).pin_memory()
        rewind()
        # The step's ~31 launches (two streams, fork / join events) as ONE CUDA graph over the static action buffer: with a
        # synchronisation per step the host's launch latency is on the critical path, and a busy host core showed up as
        # +-15 % on this number.  The same kernels run; without graph support the calls are made directly.
        graph = None
        d_act.copy_(h_act[0], non_blocking=True)
        sim.step(d_act)  # (first use of the byte-action path allocates inside the library: not during a capture)
        torch.cuda.synchronize()
        try:
            cap = torch.cuda.Stream(device=dev)
            cap.wait_stream(stream)
            with torch.cuda.stream(cap):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=cap):
                    out = sim.step(d_act)
            stream.wait_stream(cap)
            rewind()  # (the capture ran nothing, the rewind restores the window in any case)
        except Exception as ex:  # noqa: BLE001 -- fall back to direct launches
            graph = None
            res["e2e_graph_error"] = repr(ex)
            torch.cuda.synchronize()
            rewind()
        res["e2e_graph"] = graph is not None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in range(steps):
            d_act.copy_(h_act[s % n_host], non_blocking=True)
            if graph is not None:
                graph.replay()
                out = sim.out
            else:
                out = sim.step(d_act)
            h_res.copy_(out.info[:, :n_res], non_blocking=True)
            stream.synchronize()
        e1.record(stream)
        barrier()
        res["ms_e2e"] = e0.elapsed_time(e1)
        del graph
        del h_act, h_res
    # ---- pass 3: per kernel class (the events serialise nothing, but the blocking read does: its own pass)
    if per_kernel:
        rewind()
        # RoboCup: the handed-over environments' kernels on the caller's stream for this pass, so that every kernel class has a
        # duration of its own (comparable with an ncu launch list); the timed passes above run them beside the observations
        try:
            sim.debug_timing(True, serialise=True)
        except TypeError:
            sim.debug_timing(True)
        kms = {}
        n3 = min(steps, 40)
        for s in range(n3):
            sim.step(pool[s % n_pool])
            for k, (ms, n) in sim.last_kernel_ms().items():
                a = kms.setdefault(k, [0.0, 0])
                a[0] += ms
                a[1] += n
        sim.debug_timing(False)
        res["kernel_ms_per_step"] = {k: v[0] / n3 for k, v in kms.items() if v[1]}
        res["kernel_launches_per_step"] = {k: v[1] / n3 for k, v in kms.items() if v[1]}
    # ---- pass 4 (RoboCup): step + the InOutArranger kernels on the step's records (flat per-type inputs, counts, offsets)
    if arranger:
        from dynenv_b200.arranger import KernelArranger
        arr = KernelArranger.for_robocup(sim)
        rewind()
        n4 = min(steps, 20)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for s in range(n4):
            sim.step(pool[s % n_pool])
            arr.rearrange_inputs()
        a1.record(stream)
        barrier()
        res["ms_step_arranged"] = a0.elapsed_time(a1) / n4
        res["arranger_rows"] = int(arr.totals[:arr.nTypes].sum().item())
        del arr
    del sim, pool
    torch.cuda.empty_cache()
    return res


def single_env_latency(local):
    """BASELINE.json configs[0]: RoboCupEnvironment nPlayers=1, FULL observation, noise off, ONE environment -- the
    latency of a step through the batched simulator (launch-bound), over one episode"""
    import torch
    import dynenv_b200
    sim = dynenv_b200.BatchedRoboCup(1, 1, dynenv_b200.ObservationType.FULL, dynenv_b200.NoiseType.REALISTIC, 0.0, canFall=False,
                                     auto_reset=True, seed=42, device=local)
    sim.reset()
    rng = np.random.default_rng(1234)
    acts = torch.from_numpy(np.stack([rng.integers(0, 5, (240, 1, 2)), rng.integers(0, 3, (240, 1, 2)), rng.integers(0, 3, (240, 1, 2)),
                                      rng.integers(0, 7, (240, 1, 2))], -1).astype(np.float32)).cuda(local)
    for s in range(10):
        sim.step(acts[s])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(240):
        out = sim.step(acts[s])
        out.rewards.cpu()  # the caller reads the rewards back: synchronises every step
    dt = time.perf_counter() - t0
    return {"workload": "RoboCup 1v1 FULL noise off, 1 environment (configs[0])", "ms_per_step": 1e3 * dt / 240,
            "value": 240 / dt, "unit": UNIT, "steps": 240, "note": "host-synchronous step latency, 21 launches per step"}


def dropin_e2e(local):
    """the path a DynEnv user calls: make_dyn_env(...) -> vec_env.step(actions) with host numpy actions in, the reference's
    nested observation lists, rewards, dones and infos out (dynenv_b200/api.py) -- env-steps/s at three widths"""
    import dynenv_b200
    out = {}
    for n_envs, n_steps in ((8, 20), (256, 5), (4096, 2)):
        env, name = dynenv_b200.make_dyn_env(dynenv_b200.DynEnvType.ROBO_CUP, n_envs, 5, False, dynenv_b200.ObservationType.PARTIAL,
                                             dynenv_b200.NoiseType.REALISTIC, 2.0, False)
        env.reset()
        rng = np.random.default_rng(5)
        acts = [np.stack([rng.integers(0, 5, (n_envs, 10)), rng.integers(0, 3, (n_envs, 10)), rng.integers(0, 3, (n_envs, 10)),
                          rng.integers(0, 7, (n_envs, 10))], -1) for _ in range(n_steps + 1)]
        env.step(acts[0])
        t0 = time.perf_counter()
        for s in range(n_steps):
            env.step(acts[s + 1])
        dt = time.perf_counter() - t0
        out[str(n_envs)] = {"value": n_envs * n_steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / n_steps}
        env.close()
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
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
    if args.dense_obs and hasattr(wl, "sparse_obs"):
        wl.sparse_obs = False
    E = args.envs or getattr(wl, "default_envs", 65536)
    ff = FAST_FORWARD[wl.name] if args.fast_forward is None else args.fast_forward
    r = measure(wl, E, local, rank, world, args.steps, args.warmup, ff, full=args.debug_full_obs, sampler=ClockSampler(local),
                arranger=(wl.name == "robocup" and world == 1 and not args.debug_full_obs and not args.no_other))
    ms_total, ms_e2e = r["ms_total"], r["ms_e2e"]
    if world > 1:
        t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = [float(x) for x in t.tolist()]
    if rank == 0:
        peaks, which = measured_peaks()
        import ctypes
        from dynenv_b200 import _lib as _l
        tf = ctypes.c_float(0)
        fp32_peak = float(tf.value) if _l.lib().dynenv_debug_fp32_peak(local, ctypes.byref(tf)) == 0 else None
        total_envs = E * world
        step_ms = ms_total / args.steps
        value = total_envs * args.steps / (ms_total * 1e-3)
        e2e = total_envs * args.steps / (ms_e2e * 1e-3)
        P = wl.P
        # the dominant kernel class of the step and its HBM figure: algorithmic bytes per launch / launch duration
        kb = wl.kernel_bytes_per_env()
        k_ms, k_n = r["kernel_ms_per_step"], r["kernel_launches_per_step"]
        dom = max((k for k in k_ms if k in kb), key=lambda k: k_ms[k])
        dom_launch_ms = k_ms[dom] / wl.launches_per_step[dom]
        hbm_dom = kb[dom] * E / (dom_launch_ms * 1e-3) / 1e9
        step_bytes = wl.algorithmic_bytes_per_env_step()
        step_gbs = step_bytes * E / (step_ms * 1e-3) / 1e9
        tflops = FLOP_PER_ENV_STEP * E / (step_ms * 1e-3) / 1e12
        hbm_block = {"bound": "hbm", "achieved": hbm_dom, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_dom / peaks["hbm_gbs"],
                     "peak_source": which, "kernel": wl.kernel_names[dom], "kernel_ms_per_launch": dom_launch_ms,
                     "algorithmic_bytes_per_launch": kb[dom] * E,
                     "step": {"algorithmic_bytes_per_env_step": step_bytes, "achieved_gbs": step_gbs, "frac": step_gbs / peaks["hbm_gbs"]}}
        if wl.name == "image":  # the raster is HBM-write bound: its HBM figure is the primary roofline
            roofline = dict(hbm_block, traffic=None)
        else:  # the step kernels are bound by instruction issue on the FP32 / INT pipes: FP32 fraction first
            roofline = {"bound": "fp32-issue", "achieved": tflops, "peak": fp32_peak or FP32_NOMINAL_TFLOPS, "unit": "TFLOP/s",
                        "frac": tflops / (fp32_peak or FP32_NOMINAL_TFLOPS), "traffic": None,
                        "flop_per_env_step": FLOP_PER_ENV_STEP, "peak_source": "FMA probe of the library, this run" if fp32_peak else "nominal",
                        "nominal_peak_tflops": FP32_NOMINAL_TFLOPS, "hbm": hbm_block}
        roofline["kernel_ms_per_step"] = k_ms
        roofline["kernel_launches_per_step"] = k_n
        roofline["share_of_step"] = {k: v / max(sum(k_ms.values()), 1e-9) for k, v in k_ms.items()}
        roofline["note"] = ("traffic: not measurable inside the run (ncu captures of the same command: profiles/); kernel_ms_per_step from "
                            "a separate, untimed pass over the same episode window in which the handed-over environments' kernels "
                            "(phys_slow, and their share of obs) run on the caller's stream one after the other -- in the timed step "
                            "they run on a side stream beside the others' observations, so the classes add up to more than ms_per_step")
        line = {"metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict({"workload": wl.workload, "envs_per_gpu": E, "envs_total": total_envs,
                                # window of the timed steps inside the episode (240 / 600 steps long): contacts, falls and
                                # penalties build up over an episode, so the window starts mid-episode whatever --warmup says
                                "episode_steps": [ff + args.warmup, ff + args.warmup + args.steps], "fast_forward_steps": ff,
                                "l2": "state + outputs per step (%.1f GB) exceed the 126 MB L2" % (step_bytes * E / 1e9)},
                               **wl.config(L)),
                "clocks": r["clocks"], "gpu_launches": r["launches"],
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                        "cuda_graph": bool(r.get("e2e_graph", False)),
                        "note": "observations stay device-resident for the learner; rewards + done return to the host"
                                + ("; actions: MultiDiscrete indices as uint8 (dynenv_step_u8)" if getattr(wl, "byte_actions", False) else "")},
                "roofline": roofline, "diag_flags": r["diag"]}
        if "ms_step_arranged" in r:
            line["arranger"] = {"ms_per_step_with_arranger": r["ms_step_arranged"], "value": E / (r["ms_step_arranged"] * 1e-3), "unit": UNIT,
                                "rows_per_step": r["arranger_rows"],
                                "note": "step + dynenv_arrange_inputs (the learner's InOutArranger as kernels: flat per-type inputs, "
                                        "counts, offsets, maxCount; no host synchronisation), device-resident"}
        if not args.no_cpu and world == 1:
            cpu_steps = args.cpu_steps or wl.cpu_steps
            rate, threads, dt = cpu_env_steps_per_s(wl, cpu_steps)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d host threads x %d env-steps of the same workload (%.1f s), C restatement "
                                              "of the reference" % (threads, cpu_steps, dt)}
        if not args.no_other and world == 1 and wl.name == "robocup" and not args.debug_full_obs:
            # short legs of the other BASELINE.json configs, so that the driver's record carries them
            other = {}
            for name, (e_o, k_o) in (("driving", (65536, 20)), ("image", (4096, 2))):
                try:
                    w2 = WORKLOADS[name]()
                    r2 = measure(w2, e_o, local, rank, 1, k_o, 3, FAST_FORWARD[name], per_kernel=False)
                    other[name] = {"metric": w2.metric, "workload": w2.workload, "envs": e_o, "steps": k_o,
                                   "episode_steps": [FAST_FORWARD[name] + 3, FAST_FORWARD[name] + 3 + k_o],
                                   "value": e_o * k_o / (r2["ms_total"] * 1e-3), "unit": UNIT, "ms_per_step": r2["ms_total"] / k_o,
                                   "e2e": e_o * k_o / (r2["ms_e2e"] * 1e-3), "gpu_launches": r2["launches"]}
                except Exception as ex:  # a leg must not cost the headline
                    other[name] = {"error": repr(ex)}
            try:
                other["single_env"] = single_env_latency(local)
            except Exception as ex:
                other["single_env"] = {"error": repr(ex)}
            line["other_workloads"] = other
            try:
                line["e2e_dropin"] = dict(dropin_e2e(local), note="make_dyn_env('RoboCup', n, 5, ...).step(numpy actions): the reference's "
                                          "nested observation lists built on the host (dynenv_b200/api.py)")
            except Exception as ex:
                line["e2e_dropin"] = {"error": repr(ex)}
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
    ap.add_argument("--dense-obs", action="store_true", help="RoboCup: zero-filled padded observation records instead of the sparse stream")
    ap.add_argument("--no-other", action="store_true", help="skip the short legs of the other workloads / the drop-in API")
    ap.add_argument("--fast-forward", type=int, default=None, help="untimed steps in front of the warm-up (default: to mid-episode)")
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
