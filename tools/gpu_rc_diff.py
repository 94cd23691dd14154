"""GPU box: one step of a golden scenario on the GPU and in the oracle; prints where an observation list differs
usage: python tools/gpu_rc_diff.py <scenario> <list> [env ...]"""
import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import golden_cfg, load_golden, oracle_step_batch, rel_err
from test_gpu_robocup import _sim
from oracle.oracle import words_to_f32_record
name, key = sys.argv[1], sys.argv[2]
g = load_golden(name); cfg = golden_cfg(g)
words = np.concatenate([g["state0"][None], g["states"][:-1]], 0)
recs = words_to_f32_record(words)
seed = 1000 + len(name)
sim = _sim(cfg, len(recs), seed)
sim.set_state(torch.from_numpy(recs.view(np.int32)))
out = sim.step(torch.from_numpy(np.ascontiguousarray(g["actions"], dtype=np.float32)).cuda())
torch.cuda.synchronize()
outs, ow = oracle_step_batch(cfg, recs, g["actions"], seed)
P = 2 * cfg["nPlayers"]
got_all = getattr(out, key).cpu().numpy(); cnt_all = out.counts.cpu().numpy()
envs = [int(a) for a in sys.argv[3:]] or range(len(recs))
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for e in envs:
    ref, got = outs[e][key][:, :P], got_all[e]
    for t in range(ref.shape[0]):
        for a in range(P):
            if rel_err(got[t, a], ref[t, a]) > 1e-3:
                print("env %d tick %d observer %d: counts gpu %s oracle %s" % (e, t, a, cnt_all[e, t, a].tolist(), outs[e]["counts"][t, a].tolist()))
                n = int(max(cnt_all[e, t, a].max(), outs[e]["counts"][t, a].max()))
                for r in range(min(n, ref.shape[2])):
                    if rel_err(got[t, a, r], ref[t, a, r]) > 1e-3:
                        print("   row %d gpu %s\n          ref %s" % (r, got[t, a, r], ref[t, a, r]))
