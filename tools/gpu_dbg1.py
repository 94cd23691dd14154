"""Dev tool (GPU box): one substep of one golden env with the RC_DEBUG_PRINT build"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import golden_cfg, load_golden
from oracle.oracle import words_to_f32_record
import dynenv_b200
name, env, sub = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
g = load_golden(name); cfg = golden_cfg(g)
words = np.concatenate([g["state0"][None], g["states"][:-1]], 0)
recs = words_to_f32_record(words)
if len(sys.argv) > 4:
    tr = np.load(os.path.join(ROOT, "gpurun_out", sys.argv[4]))["trace"]
    recs[env] = tr[sub - 1]
n = len(recs)
keep = np.zeros(n, bool); keep[env] = True
sim = dynenv_b200.BatchedRoboCup(n, cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"],
                                 allowHeadTurn=cfg["allowHeadTurn"], randomInit=cfg["randomInit"], canFall=cfg["canFall"],
                                 useObsRewards=True, seed=1000 + len(name))
# park the other envs: no contacts -> no prints
r2 = recs.copy()
from dynenv_b200 import layout as L
sim.set_state(torch.from_numpy(r2.view(np.int32)))
sim.debug_substep_range(sub, sub + 1)
sim.step(torch.from_numpy(np.ascontiguousarray(g["actions"], dtype=np.float32)).cuda())
torch.cuda.synchronize()
