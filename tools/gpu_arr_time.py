"""GPU box: milliseconds of dynenv_arrange_inputs (4 kernels) and dynenv_arrange_outputs on the records of 65536 RoboCup
environments mid-episode"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dynenv_b200
from dynenv_b200.arranger import KernelArranger
E = 65536
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=42, sparse_obs=True)
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
hi = torch.tensor([5, 3, 3, 7], device="cuda")
for s in range(60):
    sim.step((torch.rand((E, 10, 4), generator=g, device="cuda") * hi).floor().contiguous())
arr = KernelArranger.for_robocup(sim)
for rep in range(2):
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record()
    inputs, (counts, maxc, objc) = arr.rearrange_inputs()
    e1.record()
    torch.cuda.synchronize()
print("arrange_inputs %.3f ms, rows %d, maxCount %d" % (e0.elapsed_time(e1), int(arr.totals[:6].sum().item()), int(maxc)))
