"""GPU box, library built with -DRC_DBG_COUNTERS (tools/tmp/libdbg.so): what the contact halves of the handed-over RoboCup
environments look like around step N -- active pairs per contact half, kinds of pairs, cached pairs"""
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
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=42, sparse_obs=True)
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
hi = torch.tensor([5, 3, 3, 7], device="cuda")
lib = _lib.lib()
lib.dynenv_slow_debug_counters.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 32)()
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 125):
    a = (torch.rand((E, 10, 4), generator=g, device="cuda") * hi).floor().contiguous()
    if s % 40 == 0:
        torch.cuda.synchronize(); lib.dynenv_slow_debug_counters(buf, 1)
    sim.step(a)
    if s % 40 == 0:
        torch.cuda.synchronize(); lib.dynenv_slow_debug_counters(buf, 1)
        v = list(buf)
        d = v[16:]
        n = max(d[10], 1)
        print("step %d: hand-over kernel: %.3f substeps per env-step, %.3f ordered robot phases, %.3f contact halves" % (s, v[8] / E, v[9] / E, d[10] / E))
        print("   active pairs per contact half: 0: %.3f  1: %.3f  2: %.3f  3+: %.3f ; cached pairs %.2f" % (d[0] / n, d[1] / n, d[2] / n, d[3] / n, d[9] / n))
        tot = max(sum(d[4:9]), 1)
        print("   active pairs by kind: own feet %.3f  ball-foot %.3f  feet of two robots %.3f  post-foot %.3f  ball-post %.3f" % tuple(x / tot for x in d[4:9]))
