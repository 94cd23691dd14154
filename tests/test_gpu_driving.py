"""GPU parity of the Driving path (run on the B200 box, `-m gpu`): the CUDA kernels, called through the C ABI,
against the CPU oracle (oracle/drv_oracle.c) on the same inputs and the same Philox draws: integer / flag / count
fields exactly, continuous values within 1e-4 relative (the kernels compute in float32, the oracle in float64).
Nothing here reads /root/reference."""
import numpy as np
import pytest

from helpers import drv_compare_step, drv_golden_cfg, drv_oracle_step_batch, golden_names, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-4
NAMES = golden_names("driving_")


def _sim(cfg, n, seed, **kw):
    import dynenv_b200
    return dynenv_b200.BatchedDriving(n, cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"],
                                      continuousActions=cfg["continuous"], seed=seed, **kw)


@pytest.mark.parametrize("name", NAMES)
def test_one_step_from_golden_states(name):
    """every step of every golden scenario, each started from the reference's own previous state"""
    import torch
    from oracle.oracle import drv_f32_record_to_words, drv_words_to_f32_record
    g = load_golden(name, "driving_")
    cfg = drv_golden_cfg(g)
    words = np.concatenate([g["state0"][None], g["states"][:-1]], 0)
    recs = drv_words_to_f32_record(words)
    seed = 2000 + len(name)
    sim = _sim(cfg, len(recs), seed)
    sim.set_state(torch.from_numpy(recs.view(np.int32)))
    out = sim.step(torch.from_numpy(np.ascontiguousarray(g["actions"], dtype=np.float32)).cuda())
    torch.cuda.synchronize()
    assert sim.last_launch_count() == 2  # substeps + observation/finish
    gw = drv_f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
    outs, ow = drv_oracle_step_batch(cfg, recs, g["actions"], seed)
    bad = drv_compare_step(cfg, out, gw, outs, ow, TOL)
    assert not bad, bad[:12]


@pytest.mark.parametrize("nPlayers,obs,noise,mag,cont", [
    (10, 1, 1, 2.0, False), (10, 1, 0, 3.0, True), (4, 0, 1, 0.0, False), (7, 1, 1, 5.0, True)])
def test_rollout_states(nPlayers, obs, noise, mag, cont):
    """device reset + random play; every few steps the GPU state is handed to the oracle and one step is compared"""
    import torch
    from dynenv_b200 import layout as L
    from oracle.oracle import DrvOracle, drv_f32_record_to_words
    cfg = dict(nPlayers=nPlayers, obsType=obs, noiseType=noise, noiseMag=mag, continuous=cont)
    n, seed, N = 48, 91, nPlayers
    sim = _sim(cfg, n, seed)
    out = sim.reset()
    torch.cuda.synchronize()
    st = drv_f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
    counts0 = out.counts.cpu().numpy()
    for e in range(0, n, 7):  # reset parity: the oracle draws the same scene and the same first observation
        orc = DrvOracle(nPlayers, obs, noise, mag, cont, seed=seed, env_id=e)
        orc.reset()
        w = orc.get_state()
        for name, (off, cnt, kind) in L.DRV_FIELDS.items():
            assert np.allclose(st[e, off:off + cnt], w[off:off + cnt], rtol=1e-6, atol=1e-4), (e, name)
        o = orc.reset_obs()
        assert np.array_equal(counts0[e], o["counts"][:N]), (e, counts0[e].tolist(), o["counts"][:N].tolist())
    rng = np.random.default_rng(6)
    total_bad = []
    for step in range(40):
        a = rng.uniform(-3, 3, (n, N, 2)).astype(np.float32) if cont else rng.integers(0, 3, (n, N, 2)).astype(np.float32)
        if step % 5 == 4:
            recs = sim.get_state().cpu().numpy().view(np.uint32)
            out = sim.step(torch.from_numpy(a).cuda())
            torch.cuda.synchronize()
            gw = drv_f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
            outs, ow = drv_oracle_step_batch(cfg, recs, a, seed)
            total_bad += [(step,) + b for b in drv_compare_step(cfg, out, gw, outs, ow, TOL)]
        else:
            sim.step(torch.from_numpy(a).cuda())
    assert not total_bad, total_bad[:12]  # 8 compared steps x n environments, no budget


def test_noise_statistics_match_in_distribution():
    """north_star: noise statistics match the reference in distribution (KS test, p > 0.01).  The first noisy observation
    of freshly drawn scenes over many environments (GPU, Philox streams) against the oracle fed with an independent
    sequential stream of uniforms (tape mode, the mode in which it is pinned against the reference's own Python)."""
    import torch
    from scipy import stats
    import dynenv_b200
    from oracle.oracle import DrvOracle
    n_gpu, n_cpu, N = 4096, 500, 10
    sim = dynenv_b200.BatchedDriving(n_gpu, N, 1, 1, 3.0, continuousActions=True, seed=5)
    out = sim.reset()
    torch.cuda.synchronize()
    cnt = out.counts.cpu().numpy()[:, 0]           # observer 0: rows per list (cars, obstacles, pedestrians, lanes)
    g_car = out.car.cpu().numpy()[:, 0, 0, :4]     # first car row: position, orientation
    g_obs = out.obs.cpu().numpy()[:, 0, 0, :2]
    g_ped = out.ped.cpu().numpy()[:, 0, 0, :2]
    g_lane = out.lane.cpu().numpy()[:, 0, 0, 0]
    rng = np.random.default_rng(78)
    c_cnt, c_car, c_obs, c_ped, c_lane = [], [], [], [], []
    for i in range(n_cpu):
        orc = DrvOracle(N, 1, 1, 3.0, True)
        orc.set_tape(rng.random(20000))
        orc.reset()
        o = orc.reset_obs()
        assert not orc.tape_underflow()
        c_cnt.append(o["counts"][0]); c_car.append(o["car"][0, 0, :4]); c_obs.append(o["obs"][0, 0, :2])
        c_ped.append(o["ped"][0, 0, :2]); c_lane.append(o["lane"][0, 0, 0])
    c_cnt, c_car, c_obs, c_ped, c_lane = [np.array(x) for x in (c_cnt, c_car, c_obs, c_ped, c_lane)]
    for name, k, g, c in (("car", 0, g_car, c_car), ("obs", 1, g_obs, c_obs), ("ped", 2, g_ped, c_ped)):
        for f in range(g.shape[1]):
            p = stats.ks_2samp(g[cnt[:, k] > 0, f], c[c_cnt[:, k] > 0, f]).pvalue
            assert p > 0.01, (name, f, p)
    p = stats.ks_2samp(g_lane[cnt[:, 3] > 0], c_lane[c_cnt[:, 3] > 0]).pvalue
    assert p > 0.01, ("lane", p)
    # sighting counts (false negatives, false positives, misclassification between cars and obstacles): the means agree
    # within 4 sigma of the sampling error
    for k in range(3):
        mg, mc = cnt[:, k].mean(), c_cnt[:, k].mean()
        sigma = np.sqrt(cnt[:, k].var() / n_gpu + c_cnt[:, k].var() / n_cpu) + 1e-9
        assert abs(mg - mc) < 4 * sigma + 1e-3, (k, mg, mc, sigma)


def test_overlapping_bodies_at_rest_collide_after_injection():
    """two cars put on top of each other with zero velocity and no cached pair (a hand-made state, first substep after the
    injection): the quiet-pair logic must not skip them -- the crash is detected like in the oracle"""
    import torch
    import dynenv_b200
    from dynenv_b200 import layout as L
    from oracle.oracle import drv_f32_record_to_words
    cfg = dict(nPlayers=4, obsType=1, noiseType=0, noiseMag=0.0, continuous=False)
    n, seed = 4, 17
    sim = _sim(cfg, n, seed)
    sim.reset()
    st = sim.get_state()
    F = L.DRV_FIELDS
    for name in ("cpx", "cpy", "cpxLo", "cpyLo", "cang", "cangLo"):
        st[:, F[name][0] + 1] = st[:, F[name][0] + 0]          # car 1 onto car 0
    stf = st.view(torch.float32)                               # the same storage, float view
    stf[:, F["cpx"][0] + 1] = stf[:, F["cpx"][0]] + 3.0
    for name in ("cvx", "cvy", "cw", "cvbx", "cvby", "cwb"):
        st[:, F[name][0]:F[name][0] + 10] = 0                  # everything at rest
    st[:, F["stepped"][0]] = 0
    st[:, F["nArb"][0]] = 0
    sim.set_state(st)
    recs = sim.get_state().cpu().numpy().view(np.uint32)
    a = np.ones((n, 4, 2), np.float32)                           # categorical (1, 1): no acceleration, no steering
    out = sim.step(torch.from_numpy(a).cuda())
    torch.cuda.synchronize()
    gw = drv_f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
    outs, ow = drv_oracle_step_batch(cfg, recs, a, seed)
    fl_g = gw[:, F["cflags"][0]:F["cflags"][0] + 2].astype(int)
    fl_o = ow[:, F["cflags"][0]:F["cflags"][0] + 2].astype(int)
    assert (fl_o & L.DRV_CF_CRASHED).all(), fl_o             # the oracle crashes both cars ...
    assert np.array_equal(fl_g, fl_o), (fl_g, fl_o)           # ... and so does the kernel
    bad = drv_compare_step(cfg, out, gw, outs, ow, TOL)
    assert not [b for b in bad if b[1] in ("diag", "done", "state.cflags", "rewards")], bad[:10]
