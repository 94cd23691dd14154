"""GPU box: what the cached contact pairs are (by shape kinds and by the state of the robots involved)"""
import sys, collections
import numpy as np
import torch
sys.path.insert(0, ".")
import dynenv_b200
from dynenv_b200 import layout as L
E = 16384
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
hi = torch.tensor([5, 3, 3, 7], device="cuda")
F = L.RC_FIELDS
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 221):
    a = (torch.rand((E, 10, 4), generator=g, device="cuda") * hi).floor().contiguous()
    sim.step(a)
    if s in (60, 120, 220):
        st = sim.get_state().cpu().numpy()
        narb = st[:, F["nArb"][0]]
        keys = st[:, L.C["RC_S_ARBKEY"]:L.C["RC_S_ARBKEY"] + 32]
        fl = st[:, F["rflags"][0]:F["rflags"][0] + 10]
        vx = st[:, F["vx"][0]:F["vx"][0] + 21].view(np.float32); vy = st[:, F["vy"][0]:F["vy"][0] + 21].view(np.float32)
        kinds = collections.Counter(); states = collections.Counter(); speed = collections.Counter()
        nenv = 0
        for e in np.nonzero(narb > 0)[0]:
            nenv += 1
            for k in range(narb[e]):
                key = int(keys[e, k]); a_, b_ = key & 255, (key >> 8) & 255; stt = (key >> 16) & 255
                def kind(s_): return "foot" if s_ < 20 else ("ball" if s_ == 20 else "post")
                kk = tuple(sorted((kind(a_), kind(b_))))
                if a_ < 20 and b_ < 20 and a_ // 2 == b_ // 2: kk = ("own feet",)
                kinds[kk + (("cached" if stt == 3 else "touching"),)] += 1
                for s_ in (a_, b_):
                    if s_ < 20:
                        f = int(fl[e, s_ // 2])
                        states["fallen" if f & 4 else ("penalized" if f & 8 else ("kicking" if f & 1 else "up"))] += 1
                        speed["moving" if abs(vx[e, s_]) + abs(vy[e, s_]) > 1 else "still"] += 1
        print("step %d: %d of %d envs have pairs; mean pairs %.2f" % (s, nenv, E, narb[narb > 0].mean() if nenv else 0))
        print("  kinds:", dict(kinds)); print("  robots:", dict(states), dict(speed), flush=True)
