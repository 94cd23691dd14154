"""GPU box experiment: steps a batch made only of `quiet` late-episode environments (for ncu)"""
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
n = int(sys.argv[1])
for s in range(n):
    sim.step(act())
if n > 50:
    S = sim.get_state()
    F = L.RC_FIELDS
    fl = S[:, F["rflags"][0]:F["rflags"][0] + 10]
    m = (S[:, F["nArb"][0]] == 0) & ~((fl & (1 | 4 | 8)) > 0).any(1)
    idx = m.nonzero().squeeze(1)
    sim.set_state(S[idx[torch.arange(E, device="cuda") % len(idx)]].contiguous())
sim.debug_timing(True)
sim.step(act())
print("phys ms", sim.last_kernel_ms()["phys"][0])
