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
sim = dynenv_b200.BatchedRoboCup(n, 5, dynenv_b200.ObservationType.IMAGE, 1, 2.0, allowHeadTurn=True, canFall=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(1234)
for _ in range(23):  # the bench's image workload: categorical columns + head turn uniform in [-3, 3)
    a = torch.rand((n, P, 4), generator=g, device="cuda")
    a[..., :3] = (a[..., :3] * torch.tensor([5, 3, 3], device="cuda")).floor()
    a[..., 3] = a[..., 3] * 6.0 - 3.0
    sim.step(a.contiguous())
lib = _lib.lib()
lib.dynenv_debug_raster_prof.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 24)()
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
names = ["", "setup", "conic rows", "drawing", "", "", "", "store", "total"]
for k in (1, 2, 3, 7, 8):
    print("%-32s %10.0f cycles per non-empty plane" % (names[k], v[k] / nb))
print("non-empty planes without / with the conic: %d / %d" % (v[4], v[5]))
print("big-primitive loop %.0f cycles; per non-empty plane: %.2f primitives (%.2f big), %.1f conic segments" % (v[6] / nb, v[13] / nb, v[14] / nb, v[15] / nb))
print("total cycles by channel (ball, robot, post, lines):", [round(x / nb) for x in v[9:13]])
w = v[16:24]
for k, nm in ((0, "with the conic"), (4, "without")):
    m = max(w[k], 1)
    print("lines channel %-15s %7d planes: total %7.0f cycles, big-primitive loop %6.0f, work queue %6.0f" % (nm, w[k], w[k + 1] / m, w[k + 2] / m, w[k + 3] / m))
