"""GPU box: the parity table of the RoboCup tests (one step from every golden state and the four rollouts, GPU vs oracle on
identical draws): per scenario the environments that disagree at 1e-4 and which certificate (tests/helpers.py) explains
each disagreement.  Output committed as profiles/r02_parity_table.txt."""
import sys, collections
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import golden_cfg, golden_names, load_golden, mismatch_table, oracle_step_batch
from test_gpu_robocup import STIFF, _sim, check_step
from dynenv_b200 import layout as L
from oracle.oracle import f32_record_to_words, words_to_f32_record
rows, detail = [], []


def account(name, n, res):
    bad, explained, unexplained = res
    certs = collections.Counter()
    for env, cs in collections.defaultdict(set, {}).items():
        pass
    per_env = collections.defaultdict(set)
    for b in explained:
        per_env[b[0]].add(b[-1])
    for cs in per_env.values():
        for c in cs:
            certs[c] += 1
    rows.append((name, n, len(set(b[0] for b in bad)), certs, len(set(b[0] for b in unexplained))))
    fields = collections.Counter((b[1], b[-1]) for b in explained)
    if fields or unexplained:
        detail.append("%s: " % name + ", ".join("%s[%s] x%d" % (k[0], k[1], v) for k, v in sorted(fields.items()))
                      + ("  UNEXPLAINED: %s" % [str(u)[:300] for u in unexplained[:6]] if unexplained else ""))


for name in golden_names():
    g = load_golden(name); cfg = golden_cfg(g)
    words = np.concatenate([g["state0"][None], g["states"][:-1]], 0)
    recs = words_to_f32_record(words)
    seed = 1000 + len(name)
    sim = _sim(cfg, len(recs), seed)
    sim.set_state(torch.from_numpy(recs.view(np.int32)))
    out = sim.step(torch.from_numpy(np.ascontiguousarray(g["actions"], dtype=np.float32)).cuda())
    torch.cuda.synchronize()
    gw = f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
    outs, ow = oracle_step_batch(cfg, recs, g["actions"], seed)
    account("golden " + name, len(recs), check_step(name, cfg, out, gw, outs, ow, f32_record_to_words(recs), stiff=name in STIFF))
for (nPlayers, obs, noise, mag, canFall, rinit) in [(5, 1, 1, 2.0, True, False), (5, 1, 0, 3.0, True, True), (2, 0, 1, 0.0, False, False), (3, 1, 1, 5.0, True, True)]:
    cfg = dict(nPlayers=nPlayers, obsType=obs, noiseType=noise, noiseMag=mag, canFall=canFall, allowHeadTurn=False, randomInit=rinit)
    cfg["flags"] = ((L.DYNENV_FLAG_CAN_FALL if canFall else 0) | L.DYNENV_FLAG_USE_OBS_REWARDS | (L.DYNENV_FLAG_RANDOM_INIT if rinit else 0))
    n, seed, P = 48, 77, 2 * nPlayers
    sim = _sim(cfg, n, seed); sim.reset()
    rng = np.random.default_rng(5)
    tot = [[], [], []]
    for step in range(24):
        a = np.stack([rng.integers(0, 5, (n, P)), rng.integers(0, 3, (n, P)), rng.integers(0, 3, (n, P)), rng.integers(0, 7, (n, P))], -1).astype(np.float32)
        if step % 4 == 3:
            recs = sim.get_state().cpu().numpy().view(np.uint32)
            out = sim.step(torch.from_numpy(a).cuda()); torch.cuda.synchronize()
            gw = f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
            outs, ow = oracle_step_batch(cfg, recs, a, seed)
            res = check_step("rollout", cfg, out, gw, outs, ow, f32_record_to_words(recs))
            for k in range(3):
                tot[k] += [(step * 1000 + b[0],) + b[1:] for b in res[k]]
        else:
            sim.step(torch.from_numpy(a).cuda())
    account("rollout %dv%d obs %d noise %d/%.0f fall %d rinit %d (6 x 48)" % (nPlayers, nPlayers, obs, noise, mag, canFall, rinit), 6 * n, tot)
print("RoboCup GPU-vs-oracle parity at 1e-4 (integers exact), identical Philox draws; certificates: tests/helpers.py")
print(mismatch_table(rows))
print()
print("\n".join(detail))
