"""GPU box: per-step kernel times and contact statistics of the Driving bench workload over an episode"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import dynenv_b200
from dynenv_b200 import layout as L
E = 65536
sim = dynenv_b200.BatchedDriving(E, 10, 1, 1, 2.0, continuousActions=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
sim.debug_timing(True)
F = L.DRV_FIELDS
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    a = (torch.rand((E, 10, 2), generator=g, device="cuda") * 6 - 3).contiguous()
    sim.step(a)
    if s % 5 == 4:
        k = sim.last_kernel_ms()
        st = sim.get_state()
        narb = st[:, F["nArb"][0]].float()
        fl = st[:, F["cflags"][0]:F["cflags"][0] + 10]
        key = st[:, F["arbKey"][0]:F["arbKey"][0] + 32]
        valid = torch.arange(32, device="cuda")[None, :] < st[:, F["nArb"][0]][:, None]
        a_ = key & 255; b_ = (key >> 8) & 255
        kinds = {"bld-car": (a_ < 4), "car-car": (a_ >= 4) & (a_ < 14) & (b_ < 14), "ped-car": (a_ >= 14) & (a_ < 34), "car-obs": (b_ >= 34)}
        ks = {n: float((m & valid).sum()) / E for n, m in kinds.items()}
        print("step %3d phys %.2f ms obs %.2f ms  envs with pairs %.3f  mean pairs %.2f max %d crashed %.2f finished %.2f  %s" % (
            s, k["phys"][0], k["obs"][0], float((narb > 0).float().mean()), float(narb.mean()), int(narb.max()),
            float(((fl & 2) > 0).float().mean()), float(((fl & 1) > 0).float().mean()), ks), flush=True)
