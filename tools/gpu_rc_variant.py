"""GPU box: episode-average step time of the RoboCup bench workload for the library in DYNENV_B200_LIB"""
import os, sys
import torch
sys.path.insert(0, ".")
import dynenv_b200
E = 65536
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=42, sparse_obs=bool(int(os.environ.get("SPARSE", "0"))))
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
hi = torch.tensor([5, 3, 3, 7], device="cuda")
sim.debug_timing(True)
acc = {}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for s in range(n):
    a = (torch.rand((E, 10, 4), generator=g, device="cuda") * hi).floor().contiguous()
    sim.step(a)
    k = sim.last_kernel_ms()
    b = "early" if s < 30 else ("late" if s >= n - 30 else "mid")
    for kk in ("phys", "obs", "phys_slow"):
        acc.setdefault((b, kk), []).append(k[kk][0])
    acc.setdefault(("all", "step"), []).append(k["phys"][0] + k["obs"][0] + k["finish"][0] + k["phys_slow"][0])
m = lambda v: sum(v) / len(v)
print("%-28s phys early %.2f late %.2f + handed over %.2f / %.2f | obs early %.2f late %.2f | mean step %.2f ms -> %.2f M env-steps/s" % (
    os.path.basename(os.environ.get("DYNENV_B200_LIB", "default")), m(acc[("early", "phys")]), m(acc[("late", "phys")]),
    m(acc[("early", "phys_slow")]), m(acc[("late", "phys_slow")]),
    m(acc[("early", "obs")]), m(acc[("late", "obs")]), m(acc[("all", "step")]), E / m(acc[("all", "step")]) / 1e3))
