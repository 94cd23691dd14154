"""GPU box, library built with -DRC_DBG_COUNTERS (tools/tmp/libdbg.so): per substep, how often the register-resident RoboCup
physics takes each of its exits, in rc_phys_kernel and in rc_phys_slow_kernel, around step N of the episode"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DYNENV_B200_LIB", os.path.join(ROOT, "tools", "tmp", "libdbg.so"))
import torch
import dynenv_b200
from dynenv_b200 import _lib
E = 65536
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
hi = torch.tensor([5, 3, 3, 7], device="cuda")
lib = _lib.lib()
lib.dynenv_debug_counters.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 16)()
names = ["substeps", "ordered robot phase", "own feet apart", "group hit", "cached pairs", "kept after filter", "contact exit", "-"]
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 125):
    a = (torch.rand((E, 10, 4), generator=g, device="cuda") * hi).floor().contiguous()
    if s % 40 == 0:
        torch.cuda.synchronize(); lib.dynenv_debug_counters(buf, 1)
    sim.step(a)
    if s % 40 == 0:
        torch.cuda.synchronize(); lib.dynenv_debug_counters(buf, 1)
        v = list(buf)
        print("step %d: per env-step (50 substeps):" % s)
        for k in range(7):
            print("   %-22s fast %8.4f   slow %8.4f" % (names[k], v[k] / E, v[8 + k] / E))
