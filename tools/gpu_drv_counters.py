"""GPU box, library built with -DDRV_DBG_COUNTERS: how often the Driving physics takes the contact half of a substep"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DYNENV_B200_LIB", os.path.join(ROOT, "tools", "tmp", "libdrvdbg.so"))
import torch
import dynenv_b200
from dynenv_b200 import _lib
E = 65536
sim = dynenv_b200.BatchedDriving(E, 10, 1, 1, 2.0, continuousActions=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
lib = _lib.lib()
lib.dynenv_debug_drv_counters.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 8)()
for s in range(560):
    a = (torch.rand((E, 10, 2), generator=g, device="cuda") * 6 - 3).contiguous()
    if s % 100 == 50:
        torch.cuda.synchronize(); lib.dynenv_debug_drv_counters(buf, 1)
    sim.step(a)
    if s % 100 == 50:
        torch.cuda.synchronize(); lib.dynenv_debug_drv_counters(buf, 1)
        v = list(buf)
        print("step %d: substeps with a contact half %.3f, with cached pairs %.3f; environments with a contact half in the step's launch %.3f" % (
            s, v[1] / max(v[0], 1), v[2] / max(v[0], 1), v[3] / E))
