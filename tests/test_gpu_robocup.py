"""GPU parity (run on the B200 box, `-m gpu`): the CUDA path, called through the C ABI, against
the CPU oracle on the same inputs.  Both sides draw the same Philox uniforms (include/dynenv_rng.h),
so noisy observations, falls and penalties must agree as well: integer / flag / count fields
exactly, continuous values within 1e-4 relative (north_star tolerance; the kernels compute in
float32, the oracle in float64).  Nothing here reads /root/reference."""
import numpy as np
import pytest

from helpers import compare_step, golden_cfg, golden_names, is_deterministic, load_golden, oracle_step_batch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _sim(cfg, n, seed, **kw):
    import dynenv_b200
    return dynenv_b200.BatchedRoboCup(n, cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"],
                                      allowHeadTurn=cfg["allowHeadTurn"], randomInit=cfg["randomInit"],
                                      canFall=cfg["canFall"], useObsRewards=True, seed=seed, **kw)


def _one_step(cfg, recs, actions, seed):
    import torch
    from oracle.oracle import f32_record_to_words
    sim = _sim(cfg, len(recs), seed)
    sim.set_state(torch.from_numpy(recs.view(np.int32)))
    out = sim.step(torch.from_numpy(np.ascontiguousarray(actions, dtype=np.float32)).cuda())
    torch.cuda.synchronize()
    assert sim.last_launch_count() == 1
    gw = f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
    outs, ow = oracle_step_batch(cfg, recs, actions, seed)
    return compare_step(cfg, out, gw, outs, ow, TOL)


@pytest.mark.parametrize("name", golden_names())
def test_one_step_from_golden_states(name):
    """every step of every golden scenario, each started from the reference's own previous state"""
    from oracle.oracle import words_to_f32_record
    g = load_golden(name)
    cfg = golden_cfg(g)
    words = np.concatenate([g["state0"][None], g["states"][:-1]], 0)
    recs = words_to_f32_record(words)
    bad = _one_step(cfg, recs, g["actions"], seed=1000 + len(name))
    # discontinuities (friction stop, FoV edges, noise thresholds) may flip on float32 ties
    n = len(recs)
    envs_bad = sorted(set(b[0] for b in bad))
    assert len(envs_bad) <= (0 if is_deterministic(cfg) and n < 20 else max(1, n // 20)), bad[:10]


@pytest.mark.parametrize("nPlayers,obs,noise,mag,canFall,rinit", [
    (5, 1, 1, 2.0, True, False), (5, 1, 0, 3.0, True, True), (2, 0, 1, 0.0, False, False), (3, 1, 1, 5.0, True, True)])
def test_rollout_states(nPlayers, obs, noise, mag, canFall, rinit):
    """device reset + random play; every few steps the GPU state is handed to the oracle and one
    step is compared"""
    import torch
    from dynenv_b200 import layout as L
    from oracle.oracle import RcOracle, f32_record_to_words
    cfg = dict(nPlayers=nPlayers, obsType=obs, noiseType=noise, noiseMag=mag, canFall=canFall, allowHeadTurn=False,
               randomInit=rinit)
    cfg["flags"] = ((L.DYNENV_FLAG_CAN_FALL if canFall else 0) | L.DYNENV_FLAG_USE_OBS_REWARDS
                    | (L.DYNENV_FLAG_RANDOM_INIT if rinit else 0))
    n, seed, P = 48, 77, 2 * nPlayers
    sim = _sim(cfg, n, seed)
    sim.reset()
    torch.cuda.synchronize()
    # reset parity: the oracle draws the same scene
    st = f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
    for e in range(0, n, 7):
        orc = RcOracle(nPlayers, obs, noise, mag, False, cfg["flags"], seed=seed, env_id=e)
        orc.reset()
        w = orc.get_state()
        for name, (off, cnt, kind) in L.RC_FIELDS.items():
            assert np.allclose(st[e, off:off + cnt], w[off:off + cnt], rtol=1e-6, atol=1e-4), (e, name)
    rng = np.random.default_rng(5)
    total_bad = []
    for step in range(24):
        a = np.stack([rng.integers(0, 5, (n, P)), rng.integers(0, 3, (n, P)), rng.integers(0, 3, (n, P)),
                      rng.integers(0, 7, (n, P))], -1).astype(np.float32)
        if step % 4 == 3:
            recs = sim.get_state().cpu().numpy().view(np.uint32)
            out = sim.step(torch.from_numpy(a).cuda())
            torch.cuda.synchronize()
            gw = f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
            outs, ow = oracle_step_batch(cfg, recs, a, seed)
            total_bad += [(step,) + b for b in compare_step(cfg, out, gw, outs, ow, TOL)]
        else:
            sim.step(torch.from_numpy(a).cuda())
    envs_bad = set((b[0], b[1]) for b in total_bad)
    assert len(envs_bad) <= 6 * n // 20, total_bad[:10]
