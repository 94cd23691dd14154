"""Dev tool (GPU box): rollout from device reset; dumps the state before and the outputs of one step."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import OBS_KEYS, SEEN_KEYS
import dynenv_b200
n, seed, P = 48, 77, 10
sim = dynenv_b200.BatchedRoboCup(n, 5, 1, 1, 2.0, canFall=True, seed=seed)
sim.reset(); torch.cuda.synchronize()
rng = np.random.default_rng(5)
for step in range(4):
    a = np.stack([rng.integers(0, 5, (n, P)), rng.integers(0, 3, (n, P)), rng.integers(0, 3, (n, P)), rng.integers(0, 7, (n, P))], -1).astype(np.float32)
    if step == 3:
        recs = sim.get_state().cpu().numpy().view(np.uint32)
    out = sim.step(torch.from_numpy(a).cuda())
torch.cuda.synchronize()
d = {"recs": recs, "actions": a, "state_after": sim.get_state().cpu().numpy().view(np.uint32), "seed": np.array(seed)}
for k in OBS_KEYS + SEEN_KEYS + ["rewards", "done", "state_robots", "state_ball", "recon_ball", "recon_self", "recon_others", "diag"]:
    d[k] = getattr(out, k).detach().cpu().numpy()
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_rollout.npz"), **d)
