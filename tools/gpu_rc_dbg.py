"""GPU box: counters of the RC_DBG_COUNTERS build (which substeps leave the register-resident path, and why)"""
import ctypes, os, sys
os.environ["DYNENV_B200_LIB"] = os.path.abspath("dynenv_b200/libdynenv_dbg.so")
import torch
sys.path.insert(0, ".")
import dynenv_b200
from dynenv_b200 import _lib
E = 16384
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
hi = torch.tensor([5, 3, 3, 7], device="cuda")
lb = _lib.lib()
out = (ctypes.c_ulonglong * 16)()
names = ["substeps", "ordered robots", "own feet", "group hit", "cached", "touching candidates", "slow space"]
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 121):
    a = (torch.rand((E, 10, 4), generator=g, device="cuda") * hi).floor().contiguous()
    sim.step(a)
    if s % 20 == 19:
        lb.dynenv_debug_counters(out, 1)
        n = float(out[0]) or 1.0
        print("steps %3d-%3d: " % (s - 19, s) + ", ".join("%s %.4f" % (names[k], out[k] / n) for k in range(1, 7)), flush=True)
