import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynenv_b200
d = np.load(os.path.join(ROOT, "tools", "tmp", "dump_rollout.npz"))
sim = dynenv_b200.BatchedRoboCup(48, 5, 1, 1, 2.0, canFall=True, seed=int(d["seed"]))
sim.set_state(torch.from_numpy(d["recs"].view(np.int32)))
sim.step(torch.from_numpy(d["actions"]).cuda()); torch.cuda.synchronize()
