"""GPU box: mismatches of the rollout parity tests (device reset + random play, every few steps one step against the
oracle), RoboCup and Driving: which (step, env, field) disagree at 1e-4"""
import sys, collections
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import compare_step, oracle_step_batch, drv_compare_step, drv_oracle_step_batch
from dynenv_b200 import layout as L
from oracle.oracle import f32_record_to_words, drv_f32_record_to_words
import test_gpu_robocup as TR, test_gpu_driving as TD
for (nPlayers, obs, noise, mag, canFall, rinit) in [(5, 1, 1, 2.0, True, False), (5, 1, 0, 3.0, True, True), (2, 0, 1, 0.0, False, False), (3, 1, 1, 5.0, True, True)]:
    cfg = dict(nPlayers=nPlayers, obsType=obs, noiseType=noise, noiseMag=mag, canFall=canFall, allowHeadTurn=False, randomInit=rinit)
    cfg["flags"] = ((L.DYNENV_FLAG_CAN_FALL if canFall else 0) | L.DYNENV_FLAG_USE_OBS_REWARDS | (L.DYNENV_FLAG_RANDOM_INIT if rinit else 0))
    n, seed, P = 48, 77, 2 * nPlayers
    sim = TR._sim(cfg, n, seed); sim.reset()
    rng = np.random.default_rng(5); bad = []
    for step in range(24):
        a = np.stack([rng.integers(0, 5, (n, P)), rng.integers(0, 3, (n, P)), rng.integers(0, 3, (n, P)), rng.integers(0, 7, (n, P))], -1).astype(np.float32)
        if step % 4 == 3:
            recs = sim.get_state().cpu().numpy().view(np.uint32)
            out = sim.step(torch.from_numpy(a).cuda()); torch.cuda.synchronize()
            gw = f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
            outs, ow = oracle_step_batch(cfg, recs, a, seed)
            bad += [(step,) + b for b in compare_step(cfg, out, gw, outs, ow, 1e-4)]
        else:
            sim.step(torch.from_numpy(a).cuda())
    print("robocup", (nPlayers, obs, noise, mag, canFall, rinit), "bad (step, env):", len(set((b[0], b[1]) for b in bad)), "of", 6 * n,
          dict(collections.Counter(b[2] for b in bad)))
    for b in bad[:10]: print("    ", b[0], b[1], b[2], b[3] if not isinstance(b[3], float) else "%.2e" % b[3])
for (nPlayers, obs, noise, mag, cont) in [(10, 1, 1, 2.0, False), (10, 1, 0, 3.0, True), (4, 0, 1, 0.0, False), (7, 1, 1, 5.0, True)]:
    cfg = dict(nPlayers=nPlayers, obsType=obs, noiseType=noise, noiseMag=mag, continuous=cont)
    n, seed, N = 48, 91, nPlayers
    sim = TD._sim(cfg, n, seed); sim.reset()
    rng = np.random.default_rng(6); bad = []
    for step in range(40):
        a = rng.uniform(-3, 3, (n, N, 2)).astype(np.float32) if cont else rng.integers(0, 3, (n, N, 2)).astype(np.float32)
        if step % 5 == 4:
            recs = sim.get_state().cpu().numpy().view(np.uint32)
            out = sim.step(torch.from_numpy(a).cuda()); torch.cuda.synchronize()
            gw = drv_f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
            outs, ow = drv_oracle_step_batch(cfg, recs, a, seed)
            bad += [(step,) + b for b in drv_compare_step(cfg, out, gw, outs, ow, 1e-4)]
        else:
            sim.step(torch.from_numpy(a).cuda())
    print("driving", (nPlayers, obs, noise, mag, cont), "bad (step, env):", len(set((b[0], b[1]) for b in bad)), "of", 8 * n,
          dict(collections.Counter(b[2] for b in bad)))
    for b in bad[:10]: print("    ", b[0], b[1], b[2], b[3] if not isinstance(b[3], float) else "%.2e" % b[3])
