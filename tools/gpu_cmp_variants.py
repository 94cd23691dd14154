"""GPU box: state records after N steps by two library builds (subprocesses), compared field by field"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynenv_b200 import layout as L
RUN = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import dynenv_b200
E, steps = int(sys.argv[2]), int(sys.argv[3])
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=77)
sim.reset()
g = torch.Generator().manual_seed(3)
hi = torch.tensor([5, 3, 3, 7])
for s in range(steps):
    out = sim.step((torch.rand((E, 10, 4), generator=g) * hi).floor().cuda())
torch.cuda.synchronize()
np.save(sys.argv[1], sim.get_state().cpu().numpy())
""" % ROOT
libs = sys.argv[1:3]
E = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
for steps in (1, 2, 5, 10, 20, 40, 60):
    res = []
    for i, lib in enumerate(libs):
        dst = "/tmp/cmp_%d.npy" % i
        subprocess.run([sys.executable, "-c", RUN, dst, str(E), str(steps)], env=dict(os.environ, DYNENV_B200_LIB=os.path.abspath(lib)), check=True)
        res.append(np.load(dst))
    a, b = res
    bad = np.argwhere((a != b).any(1)).ravel()
    print("steps %d: %d of %d environments differ" % (steps, len(bad), E))
    if len(bad):
        e = bad[0]
        for name, (off, cnt, kind) in L.RC_FIELDS.items():
            x, y = a[e, off:off + cnt], b[e, off:off + cnt]
            if not np.array_equal(x, y):
                if kind == "i":
                    print("   env %d %-16s int  %s | %s" % (e, name, x[x != y][:6].tolist(), y[x != y][:6].tolist()))
                else:
                    xf, yf = x.view(np.float32), y.view(np.float32)
                    m = x != y
                    print("   env %d %-16s f32  max |d| %.3e  ulps %s  vals %s | %s" % (e, name, np.abs(xf[m] - yf[m]).max(), np.abs(x[m].astype(np.int64) - y[m].astype(np.int64))[:6].tolist(), xf[m][:3].tolist(), yf[m][:3].tolist()))
        break
