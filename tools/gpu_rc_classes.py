"""GPU box experiment: cost of one RoboCup step when the batch consists only of environments of one class (taken from
the state at step 200 of the bench workload): is the late-episode slowdown carried by the environments with cached
contact pairs / fallen / penalised robots, or by all of them?"""
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
act = lambda: (torch.rand((E, 10, 4), generator=g, device="cuda") * hi).floor().contiguous()
for s in range(200):
    sim.step(act())
S = sim.get_state()
F = L.RC_FIELDS
narb = S[:, F["nArb"][0]]
fl = S[:, F["rflags"][0]:F["rflags"][0] + 10]
fallen = ((fl & 4) > 0).any(1)
penal = ((fl & 8) > 0).any(1)
kick = ((fl & 1) > 0).any(1)
classes = {"all": torch.ones(E, dtype=torch.bool, device="cuda"), "pairs": narb > 0, "no pairs": narb == 0,
           "fallen": fallen, "penalised": penal, "kicking": kick & (narb == 0) & ~fallen & ~penal,
           "quiet (no pairs / fallen / penalised / kicking)": (narb == 0) & ~fallen & ~penal & ~kick}
sim.debug_timing(True)
a = act()
for name, m in classes.items():
    idx = m.nonzero().squeeze(1)
    if len(idx) == 0:
        continue
    sel = idx[torch.arange(E, device="cuda") % len(idx)]
    ph, ob = [], []
    for rep in range(3):
        sim.set_state(S[sel].contiguous())
        sim.step(a)
        k = sim.last_kernel_ms()
        ph.append(k["phys"][0]); ob.append(k["obs"][0])
    print("%-52s %5.1f%% of envs  phys %.2f ms  obs %.2f ms" % (name, 100.0 * len(idx) / E, min(ph), min(ob)), flush=True)
