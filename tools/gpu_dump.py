"""Dev tool (GPU box): one GPU step from every golden state, results dumped to gpurun_out/dump_<name>.npz
so that they can be compared with the oracle offline (tools/cmp_dump.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import FULL_KEYS, OBS_KEYS, SEEN_KEYS, golden_cfg, golden_names, load_golden  # noqa: E402
from oracle.oracle import words_to_f32_record  # noqa: E402
import dynenv_b200  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for name in (sys.argv[1:] or golden_names()):
    g = load_golden(name)
    cfg = golden_cfg(g)
    words = np.concatenate([g["state0"][None], g["states"][:-1]], 0)
    recs = words_to_f32_record(words)
    seed = 1000 + len(name)
    sim = dynenv_b200.BatchedRoboCup(len(recs), cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"],
                                     allowHeadTurn=cfg["allowHeadTurn"], randomInit=cfg["randomInit"],
                                     canFall=cfg["canFall"], useObsRewards=True, seed=seed)
    sim.set_state(torch.from_numpy(recs.view(np.int32)))
    out = sim.step(torch.from_numpy(np.ascontiguousarray(g["actions"], dtype=np.float32)).cuda())
    torch.cuda.synchronize()
    d = {"state_after": sim.get_state().cpu().numpy().view(np.uint32), "recs": recs, "seed": np.array(seed)}
    keys = (FULL_KEYS if cfg["obsType"] == 0 else OBS_KEYS + SEEN_KEYS) + ["rewards", "done", "state_robots", "state_ball",
                                                                           "recon_ball", "recon_self", "recon_others", "diag"]
    for k in keys:
        d[k] = getattr(out, k).detach().cpu().numpy()
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_%s.npz" % name), **d)
    print(name, "ok")
