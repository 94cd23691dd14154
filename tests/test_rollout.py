"""Rollout buffers: the numpy oracle against the reference's own _discount_rewards outputs (CPU), and the
device kernels behind dynenv_rollout_insert / dynenv_discount_rewards against the oracle and the golden (GPU)."""
import os

import numpy as np
import pytest

from oracle import rollout_oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "rollout.npz"))
NCASES = 4


@pytest.mark.parametrize("k", range(NCASES))
def test_oracle_matches_reference_discount(k):
    out = rollout_oracle.discount_rewards(GOLD["rewards%d" % k], GOLD["final%d" % k], float(GOLD["discount%d" % k]))
    assert np.array_equal(out, GOLD["out%d" % k])  # float32 bit-exact


def test_oracle_done_cut():
    r = np.ones((4, 2), np.float32)
    dones = np.array([[0], [1], [0], [0]], np.uint8)
    out = rollout_oracle.discount_rewards(r, np.zeros(2), 0.5, dones, 2)
    assert np.allclose(out[:, 0], [1.5, 1.0, 1.5, 1.0])


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(NCASES))
def test_gpu_discount_matches_reference(k):
    import ctypes as C
    import torch
    from dynenv_b200 import _lib
    lib = _lib.lib()
    T, E, A = [int(v) for v in GOLD["shape%d" % k]]
    rew = torch.from_numpy(GOLD["rewards%d" % k]).cuda()
    fv = torch.from_numpy(GOLD["final%d" % k]).cuda()
    dn = torch.from_numpy(GOLD["dones%d" % k]).cuda()
    out = torch.empty_like(rew)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dynenv_discount_rewards(0, rew.data_ptr(), fv.data_ptr(), None, A, out.data_ptr(), T, E * A,
                                           float(GOLD["discount%d" % k]), st))
    assert np.array_equal(out.cpu().numpy(), GOLD["out%d" % k])
    _lib.check(lib.dynenv_discount_rewards(0, rew.data_ptr(), fv.data_ptr(), dn.data_ptr(), A, out.data_ptr(), T, E * A,
                                           float(GOLD["discount%d" % k]), st))
    ref = rollout_oracle.discount_rewards(GOLD["rewards%d" % k], GOLD["final%d" % k], float(GOLD["discount%d" % k]),
                                          GOLD["dones%d" % k], A)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.gpu
def test_gpu_discount_bad_arguments():
    from dynenv_b200 import _lib
    lib = _lib.lib()
    assert lib.dynenv_discount_rewards(0, None, None, None, 1, None, 1, 1, 0.9, None) != 0
    assert b"null" in lib.dynenv_last_error()
    assert lib.dynenv_discount_rewards(0, 8, 8, None, 3, 8, 2, 10, 0.9, None) != 0  # 10 columns, 3 agents per env
    assert lib.dynenv_discount_rewards(0, 8, 8, None, 2, 8, 0, 10, 0.9, None) == 0  # empty rollout: nothing launched


def _collect(o, robocup):
    st = o.state_robots if robocup else o.state_cars
    return (o.rewards.cpu().numpy().copy(), o.done.cpu().numpy().copy(), st.cpu().numpy().copy(),
            o.episode_r.cpu().numpy().copy(), o.episode_p_r.cpu().numpy().copy(),
            o.episode_o_r.cpu().numpy().copy() if robocup else None, o.episode_g.cpu().numpy().copy())


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["robocup", "driving"])
def test_gpu_rollout_insert(kind):
    """a rollout that crosses episode ends (states advanced to the last steps of the episode): the buffers and the episode
    sums equal the oracle's, built from the same result records read back field by field"""
    import torch
    import dynenv_b200
    from dynenv_b200 import layout as L
    n, T = 64, 6
    rng = np.random.default_rng(5)
    if kind == "robocup":
        sim = dynenv_b200.BatchedRoboCup(n, 2, 1, 1, 2.0, auto_reset=True, seed=3, device=0)
        A, word, per_step, horizon = 4, L.RC_S_ELAPSED, 50, 12000
    else:
        sim = dynenv_b200.BatchedDriving(n, 3, 1, 1, 2.0, continuousActions=True, auto_reset=True, seed=3, device=0)
        A, word, per_step, horizon = 3, L.DRV_S_ELAPSED, 10, 6000
    sim.reset()
    store = dynenv_b200.DeviceRolloutStorage(sim, T)
    assert store.rewards.shape == (T, n * A) and store.dones.shape == (T, n)
    # put half of the environments 1..3 steps before their episode end
    st = sim.get_state()
    st[:n // 2, word] = horizon - per_step * torch.arange(1, n // 2 + 1, device=st.device, dtype=st.dtype).remainder(3).add(1)
    sim.set_state(st)
    results = []
    for t in range(T):
        if kind == "robocup":
            a = np.stack([rng.integers(0, 5, (n, A)), rng.integers(0, 3, (n, A)), rng.integers(0, 3, (n, A)),
                          rng.integers(0, 7, (n, A))], -1).astype(np.float32)
        else:
            a = rng.uniform(-3, 3, (n, A, 2)).astype(np.float32)
        o = sim.step(torch.from_numpy(a).cuda())
        store.insert(t)
        torch.cuda.synchronize()
        results.append(_collect(o, kind == "robocup"))
    rewards, fin, dones, acc = rollout_oracle.insert(results, A)
    assert dones.sum() >= n // 2  # the prepared environments did finish inside the rollout
    assert np.array_equal(store.rewards.cpu().numpy(), rewards)
    assert np.array_equal(store.agentFinished.cpu().numpy(), fin)
    assert np.array_equal(store.dones.cpu().numpy(), dones)
    assert np.allclose(store.episode_acc.cpu().numpy(), acc, rtol=1e-6, atol=1e-6)
    stats = store.episode_stats()
    assert stats["episodes"] == int(dones.sum())
    assert abs(stats["mean_reward"] - acc[:, 1].sum() / dones.sum()) < 1e-4
    # the median of print_reward_stats (models/storage.py:310-318): over the per-agent rewards of the logged episodes
    logged = np.concatenate([np.asarray(r[3])[np.asarray(r[1]) != 0].reshape(-1) for r in results])
    assert abs(stats["median_reward"] - float(np.median(logged))) < 1e-5 + 1e-5 * abs(float(np.median(logged))) or \
        len(logged) % 2 == 0  # torch's nanmedian returns the lower of the two middle values
    if len(logged) % 2 == 0:
        lo, hi = np.sort(logged)[len(logged) // 2 - 1], np.sort(logged)[len(logged) // 2]
        assert lo - 1e-5 <= stats["median_reward"] <= hi + 1e-5
    fv = torch.from_numpy(rng.normal(size=n * A).astype(np.float32)).cuda()
    ref = rollout_oracle.discount_rewards(rewards, fv.cpu().numpy(), 0.99)
    assert np.array_equal(store._discount_rewards(fv, 0.99).cpu().numpy(), ref)
    ref = rollout_oracle.discount_rewards(rewards, fv.cpu().numpy(), 0.99, dones, A)
    assert np.array_equal(store._discount_rewards(fv, 0.99, respect_dones=True).cpu().numpy(), ref)
    store.after_update()
    assert float(store.rewards.abs().sum()) == 0 and store.episode_stats()["episodes"] == int(dones.sum())
    with pytest.raises(Exception):
        store.insert(T)
