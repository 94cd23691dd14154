"""GPU box: per-step kernel times of the RoboCup bench workload over a whole episode (240 steps)"""
import sys
import torch
sys.path.insert(0, ".")
import dynenv_b200
from dynenv_b200 import layout as L
E = 65536
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
hi = torch.tensor([5, 3, 3, 7], device="cuda")
sim.debug_timing(True)
F = L.RC_FIELDS
tot = 0.0
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 250):
    a = (torch.rand((E, 10, 4), generator=g, device="cuda") * hi).floor().contiguous()
    sim.step(a)
    k = sim.last_kernel_ms()
    tot += k["phys"][0] + k["obs"][0] + k["finish"][0]
    if s % 10 == 9:
        st = sim.get_state()
        narb = st[:, F["nArb"][0]].float()
        fl = st[:, F["rflags"][0]:F["rflags"][0] + 10]
        print("step %3d phys %.2f obs %.2f fin %.2f ms | envs with pairs %.3f mean pairs %.2f fallen %.3f penalized %.3f | mean so far %.2f ms/step" % (
            s, k["phys"][0], k["obs"][0], k["finish"][0], float((narb > 0).float().mean()), float(narb.mean()),
            float(((fl & 4) > 0).float().mean()), float(((fl & 8) > 0).float().mean()), tot / (s + 1)), flush=True)
