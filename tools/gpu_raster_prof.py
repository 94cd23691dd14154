"""Dev experiment (GPU box): per-phase cycle counts of the raster kernel from a library built with -DRASTER_PROFILE
(tools/tmp/libprof.so): python tools/gpu_raster_prof.py"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DYNENV_B200_LIB", os.path.join(ROOT, "tools", "tmp", "libprof.so"))
import torch
import dynenv_b200
from dynenv_b200 import _lib

n, P = 1024, 10
sim = dynenv_b200.BatchedRoboCup(n, 5, dynenv_b200.ObservationType.IMAGE, 1, 2.0, allowHeadTurn=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(1234)
for _ in range(5):
    a = (torch.rand((n, P, 4), generator=g, device="cuda") * torch.tensor([5, 3, 3, 7], device="cuda")).floor()
    sim.step(a)
lib = _lib.lib()
lib.dynenv_debug_raster_prof.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 16)()
out = torch.empty((256, 5, P, 8, 480, 20), dtype=torch.int32, device="cuda")
sim.render_images(0, 256, out)
lib.dynenv_debug_raster_prof(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for c in range(4):
    sim.render_images(256 * c, 256 * (c + 1), out)
e1.record()
torch.cuda.synchronize()
lib.dynenv_debug_raster_prof(buf, 1)
v = list(buf)
nb = max(v[0], 1)
print("ms per 1024 envs %.2f; non-empty planes %d of %d" % (e0.elapsed_time(e1), v[0], n * 5 * P * 8))
names = ["", "setup", "conic rows", "prims (sum over warps / 8)", "prims (max over warps)", "conic segments (sum / 8)", "conic segments (max)", "store", "total"]
for k in (1, 2, 3, 4, 5, 6, 7, 8):
    div = 8.0 if k in (3, 4, 5, 6) else 1.0
    print("%-32s %10.0f cycles per non-empty plane" % (names[k], v[k] / div / nb))
print("total cycles by channel (ball, robot, post, lines):", [round(x / nb) for x in v[9:13]])
