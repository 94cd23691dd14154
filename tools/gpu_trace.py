"""Dev tool (GPU box): substep-by-substep states of chosen golden steps: gpurun_out/trace_<name>_<env>.npz
usage: gpu_trace.py name env [name env ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden_cfg, load_golden  # noqa: E402
from oracle.oracle import words_to_f32_record  # noqa: E402
import dynenv_b200  # noqa: E402

args = sys.argv[1:]
for name, env in zip(args[0::2], args[1::2]):
    env = int(env)
    g = load_golden(name)
    cfg = golden_cfg(g)
    words = np.concatenate([g["state0"][None], g["states"][:-1]], 0)
    recs = words_to_f32_record(words)
    seed = 1000 + len(name)
    # same env ids as the full-batch run: env `env` of a batch of len(recs)
    sim = dynenv_b200.BatchedRoboCup(len(recs), cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"],
                                     allowHeadTurn=cfg["allowHeadTurn"], randomInit=cfg["randomInit"],
                                     canFall=cfg["canFall"], useObsRewards=True, seed=seed)
    sim.set_state(torch.from_numpy(recs.view(np.int32)))
    act = torch.from_numpy(np.ascontiguousarray(g["actions"], dtype=np.float32)).cuda()
    trace = []
    for i in range(50):
        sim.debug_substep_range(i, i + 1)
        sim.step(act)
        torch.cuda.synchronize()
        trace.append(sim.get_state().cpu().numpy().view(np.uint32)[env].copy())
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "trace_%s_%d.npz" % (name, env)), trace=np.array(trace),
                        rec0=recs[env], seed=np.array(seed))
    print(name, env, "ok")
