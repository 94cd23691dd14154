"""GPU box: mean step time of the Driving bench workload over the first 150 steps for the library in DYNENV_B200_LIB"""
import os, sys
import torch
sys.path.insert(0, ".")
import dynenv_b200
E = 65536
sim = dynenv_b200.BatchedDriving(E, 10, 1, 1, 2.0, continuousActions=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
sim.debug_timing(True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
ph, ob = [], []
for s in range(n):
    a = (torch.rand((E, 10, 2), generator=g, device="cuda") * 6 - 3).contiguous()
    sim.step(a)
    k = sim.last_kernel_ms()
    ph.append(k["phys"][0]); ob.append(k["obs"][0])
m = lambda v: sum(v) / len(v)
print("%-24s phys early %.2f late %.2f | obs %.2f | mean step %.2f ms -> %.2f M env-steps/s" % (
    os.path.basename(os.environ.get("DYNENV_B200_LIB", "default")), m(ph[:20]), m(ph[-20:]), m(ob), m(ph) + m(ob), E / (m(ph) + m(ob)) / 1e3))
