"""GPU box: how many environments take the record-based half of the substep (cached pairs, broad-phase group hits, own
feet apart) at a few points of the episode"""
import sys
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
def fld(st, name, n):
    return st[:, F[name][0]:F[name][0] + n].view(torch.float32) if False else st[:, F[name][0]:F[name][0] + n]
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 121):
    a = (torch.rand((E, 10, 4), generator=g, device="cuda") * hi).floor().contiguous()
    sim.step(a)
    if s % 20 == 0:
        st = sim.get_state()
        f = st.view(torch.float32)
        px = f[:, F["px"][0]:F["px"][0] + 21]; py = f[:, F["py"][0]:F["py"][0] + 21]
        ang = f[:, F["angle"][0]:F["angle"][0] + 20] if "angle" in F else None
        narb = st[:, F["nArb"][0]]
        gx = torch.cat([(px[:, 0:20:2] + px[:, 1:20:2]) / 2, px[:, 20:21]], 1)
        gy = torch.cat([(py[:, 0:20:2] + py[:, 1:20:2]) / 2, py[:, 20:21]], 1)
        gr = torch.tensor([17.5] * 10 + [10.0], device="cuda")
        d = ((gx[:, :, None] - gx[:, None, :]) ** 2 + (gy[:, :, None] - gy[:, None, :]) ** 2).sqrt()
        rr = gr[None, :, None] + gr[None, None, :] + 0.05
        hit = (d < rr) & ~torch.eye(11, dtype=torch.bool, device="cuda")[None]
        posts = torch.tensor([[70., 450.], [70., 290.], [970., 450.], [970., 290.]], device="cuda")
        dp = ((gx[:, :, None] - posts[None, None, :, 0]) ** 2 + (gy[:, :, None] - posts[None, None, :, 1]) ** 2).sqrt()
        hitp = dp < (gr[None, :, None] + 10.05)
        apart = ((px[:, 0:20:2] - px[:, 1:20:2]).abs() + (py[:, 0:20:2] - py[:, 1:20:2]).abs()) >= 2
        fl = st[:, F["rflags"][0]:F["rflags"][0] + 10]
        print("step %3d: cached pairs %.3f | group hit %.3f (robot-robot %.3f, ball %.3f, post %.3f) | feet apart %.3f | any %.3f | kicking %.3f" % (
            s, float((narb > 0).float().mean()), float((hit.any(2).any(1) | hitp.any(2).any(1)).float().mean()),
            float(hit[:, :10, :10].any(2).any(1).float().mean()), float(hit[:, 10, :].any(1).float().mean()),
            float(hitp.any(2).any(1).float().mean()), float(apart.any(1).float().mean()),
            float(((narb > 0) | hit.any(2).any(1) | hitp.any(2).any(1) | apart.any(1)).float().mean()),
            float(((fl & 1) > 0).any(1).float().mean())), flush=True)
