"""Dev check (GPU box): is the Driving "quiet pair" skip exact?  `save`: runs 4096 environments for 160 steps with the
library in DYNENV_B200_LIB and saves the state after every step to /tmp; `compare`: the same run with another library
(e.g. the -DDRV_NO_QUIET_SKIP build tools/tmp/libnoskip.so), reporting the first differing step / environment / fields."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import dynenv_b200
from dynenv_b200 import layout as L

mode = sys.argv[1]
E, STEPS = 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 160
sim = dynenv_b200.BatchedDriving(E, 10, 1, 1, 2.0, continuousActions=True, auto_reset=True, seed=7)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(99)
reported = 0
for s in range(STEPS):
    a = torch.rand((E, 10, 2), generator=g, device="cuda") * 6 - 3
    sim.step(a)
    st = sim.get_state().cpu().numpy()
    if mode == "save":
        np.save("/tmp/qdump_%d.npy" % s, st)
        continue
    ref = np.load("/tmp/qdump_%d.npy" % s)
    bad = np.nonzero((st != ref).any(1))[0]
    if len(bad) and reported < 3:
        reported += 1
        e = int(bad[0])
        print("step %d: %d environments differ; env %d:" % (s, len(bad), e))
        for name, (off, cnt, kind) in L.DRV_FIELDS.items():
            x, y = st[e, off:off + cnt], ref[e, off:off + cnt]
            if (x != y).any():
                idx = np.nonzero(x != y)[0][:6]
                xv = x.view(np.float32) if kind == "f" else x
                yv = y.view(np.float32) if kind == "f" else y
                print("   %-16s idx %s this %s saved %s" % (name, idx.tolist(), xv[idx].tolist(), yv[idx].tolist()))
        fl = st[e, L.DRV_FIELDS["cflags"][0]:L.DRV_FIELDS["cflags"][0] + 10]
        print("   nArb %d / %d, cflags %s" % (st[e, L.DRV_FIELDS["nArb"][0]], ref[e, L.DRV_FIELDS["nArb"][0]], fl.tolist()))
        keys = st[e, L.DRV_FIELDS["arbKey"][0]:L.DRV_FIELDS["arbKey"][0] + 8]
        print("   arbKey (this)  %s" % [hex(int(k) & 0xffffffff) for k in keys])
        keys = ref[e, L.DRV_FIELDS["arbKey"][0]:L.DRV_FIELDS["arbKey"][0] + 8]
        print("   arbKey (saved) %s" % [hex(int(k) & 0xffffffff) for k in keys])
print("done", mode, "differences reported:", reported)
