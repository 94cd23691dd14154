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


# Every field of every environment is compared: integer / flag / count fields exactly, floats at 1e-4 relative.  A
# disagreement is accepted only with a CERTIFICATE that it is a float32 tie at one of the reference's hard thresholds
# (tests/helpers.py: on_line, argmin_tie), computed from the oracle's own state; hand-made deep inter-penetration scenes
# compare their continuous state at STIFF_TOL (impulses of ~1e7 through ten Gauss-Seidel sweeps), everything discrete
# exactly.  tools/gpu_rc_mismatch.py prints the table of what actually disagrees (profiles/r02_parity_table.txt).
STIFF = {"ball_at_foot", "face_to_face", "face_to_face_fall", "scrum", "scrum_fall"}


def check_step(name, cfg, out, gw, outs, ow, words_before, stiff=False):
    """the parity rule of this file; returns (mismatching environments, explained, unexplained)"""
    from helpers import STIFF_TOL, certify
    P = 2 * cfg["nPlayers"]
    bad = compare_step(cfg, out, gw, outs, ow, TOL)
    bad_stiff = []
    if stiff:  # continuous state of the inter-penetration scenes at STIFF_TOL; what still differs there must be certified
        loose = compare_step(cfg, out, gw, outs, ow, STIFF_TOL, line_tol=2 * STIFF_TOL)
        keep = set((b[0], b[1]) for b in loose)  # (integer fields compare exactly at any tolerance: they stay)
        bad_stiff = [b for b in bad if (b[0], b[1]) not in keep]
        bad = loose
    explained, unexplained = certify(bad, words_before, ow, P, outs, out)
    return bad + bad_stiff, explained + [b + ("stiff",) for b in bad_stiff], unexplained


@pytest.mark.parametrize("name", golden_names())
def test_one_step_from_golden_states(name):
    """every step of every golden scenario, each started from the reference's own previous state"""
    from oracle.oracle import f32_record_to_words, words_to_f32_record
    import torch
    g = load_golden(name)
    cfg = golden_cfg(g)
    words = np.concatenate([g["state0"][None], g["states"][:-1]], 0)
    recs = words_to_f32_record(words)
    seed = 1000 + len(name)
    sim = _sim(cfg, len(recs), seed)
    sim.set_state(torch.from_numpy(recs.view(np.int32)))
    out = sim.step(torch.from_numpy(np.ascontiguousarray(g["actions"], dtype=np.float32)).cuda())
    torch.cuda.synchronize()
    # 5 x (substep block, its handed-over environments, observation tick of both groups [PARTIAL: lines + points]) + finish
    assert sim.last_launch_count() == (31 if cfg["obsType"] == 1 else 21)  # 1 = DYNENV_OBS_PARTIAL
    gw = f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
    outs, ow = oracle_step_batch(cfg, recs, g["actions"], seed)
    bad, explained, unexplained = check_step(name, cfg, out, gw, outs, ow, f32_record_to_words(recs), stiff=name in STIFF)
    assert not unexplained, unexplained[:10]


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
            total_bad += [(step,) + b for b in check_step("rollout", cfg, out, gw, outs, ow, f32_record_to_words(recs))[2]]
        else:
            sim.step(torch.from_numpy(a).cuda())
    assert not total_bad, total_bad[:10]  # no budget: every disagreement needs a certificate (check_step)


def test_noise_statistics_match_in_distribution():
    """north_star: noise statistics match the reference in distribution (KS test, p > 0.01).  The kernels draw from
    Philox streams, the reference from a sequential generator: the noisy first observation of the same scene over
    many environments (GPU, Philox) against the oracle fed with an independent sequential stream of uniforms (tape
    mode, the mode in which it is pinned against the reference's own Python)."""
    import torch
    from scipy import stats
    import dynenv_b200
    from dynenv_b200 import layout as L
    from oracle.oracle import RcOracle
    n_gpu, n_cpu = 4096, 600
    sim = dynenv_b200.BatchedRoboCup(n_gpu, 2, 1, 1, 3.0, canFall=True, seed=5)
    out = sim.reset()
    torch.cuda.synchronize()
    g_ball = out.ball.cpu().numpy()[:, 0, 0, 0, :3]        # first ball row of observer 0, tick 0: x, y, size
    g_nb = out.counts.cpu().numpy()[:, 0, 0, 0]
    g_line = out.line.cpu().numpy()[:, 0, 0, 0, 0]          # first line row: distance
    g_nl = out.counts.cpu().numpy()[:, 0, 0, 5]
    rng = np.random.default_rng(77)
    c_ball, c_nb, c_line, c_nl = [], [], [], []
    flags = L.DYNENV_FLAG_CAN_FALL | L.DYNENV_FLAG_USE_OBS_REWARDS
    for i in range(n_cpu):
        orc = RcOracle(2, 1, 1, 3.0, False, flags)
        orc.set_tape(rng.random(4000))
        orc.reset()
        o = orc.reset_obs()
        assert not orc.tape_underflow()
        c_ball.append(o["ball"][0, 0, 0, :3]); c_nb.append(o["counts"][0, 0, 0])
        c_line.append(o["line"][0, 0, 0, 0]); c_nl.append(o["counts"][0, 0, 5])
    c_ball, c_nb, c_line, c_nl = np.array(c_ball), np.array(c_nb), np.array(c_line), np.array(c_nl)
    for k in range(3):  # position noise and size noise of the ball sighting (where it was sighted)
        p = stats.ks_2samp(g_ball[g_nb > 0, k], c_ball[c_nb > 0, k]).pvalue
        assert p > 0.01, ("ball", k, p)
    p = stats.ks_2samp(g_line[g_nl > 0], c_line[c_nl > 0]).pvalue
    assert p > 0.01, ("line", p)
    # false negatives: the sighting rates agree within 4 sigma of the binomial error
    for g, c in ((g_nb > 0, c_nb > 0), (g_nl == g_nl.max(), c_nl == g_nl.max())):
        pg, pc = g.mean(), c.mean()
        sigma = np.sqrt(pc * (1 - pc) / n_cpu + pg * (1 - pg) / n_gpu) + 1e-9
        assert abs(pg - pc) < 4 * sigma + 1e-3, (pg, pc, sigma)


def test_noise_statistics_every_list_and_false_positives():
    """The same comparison for everything else the noise model touches (cutils.py:417-468, RoboCupEnvironment.py:1254-1316):
    how many rows every list has per observer -- false negatives, misclassified balls / crosses changing lists, random
    false positives per class, false balls next to robots -- as a chi-square test of homogeneity of the count histograms;
    position / size noise of robots, posts, crosses and line crosses and the angle noise of line crosses as two-sample KS
    tests.  GPU: Philox streams over 8192 environments; oracle: an independent sequential stream for each of 1500.  The
    family of tests is held to 0.01 overall (Bonferroni: every p-value above 0.01 / number of tests)."""
    import torch
    from scipy import stats
    import dynenv_b200
    from dynenv_b200 import layout as L
    from oracle.oracle import RcOracle
    n_gpu, n_cpu, P = 8192, 1500, 4
    sim = dynenv_b200.BatchedRoboCup(n_gpu, 2, 1, 1, 3.0, canFall=True, seed=6)
    out = sim.reset()
    torch.cuda.synchronize()
    names = ["ball", "robot", "post", "cross", "fcross", "line"]
    G = {k: getattr(out, k).cpu().numpy() for k in names}          # [n, T, P, cap, F]
    G["counts"] = out.counts.cpu().numpy()                          # [n, T, P, 6]
    rng = np.random.default_rng(78)
    flags = L.DYNENV_FLAG_CAN_FALL | L.DYNENV_FLAG_USE_OBS_REWARDS
    C = {k: [] for k in names + ["counts"]}
    for i in range(n_cpu):
        orc = RcOracle(2, 1, 1, 3.0, False, flags)
        orc.set_tape(rng.random(6000))
        orc.reset()
        o = orc.reset_obs()
        assert not orc.tape_underflow()
        for k in names:
            C[k].append(np.array(o[k][:, :P]))
        C["counts"].append(np.array(o["counts"][:, :P]))
    C = {k: np.array(v) for k, v in C.items()}
    pvals = {}
    # One sample per environment and test: the scene itself is drawn per environment (jittered kick-off spots, a
    # permutation of who takes which), so the five ticks and the four observers of ONE environment are dependent samples
    # and must not be pooled.  List lengths per (observer, type) at tick 0:
    for a in range(P):
        for k in range(6):
            gc, cc = G["counts"][:, 0, a, k], C["counts"][:, 0, a, k]
            hi = int(max(gc.max(), cc.max())) + 1
            hg, hc = np.bincount(gc, minlength=hi).astype(float), np.bincount(cc, minlength=hi).astype(float)
            # merge sparse bins from the top so that every expected cell count is >= 5
            while len(hg) > 1 and (hg[-1] + hc[-1]) * min(hg.sum(), hc.sum()) / (hg.sum() + hc.sum()) < 5:
                hg[-2] += hg[-1]; hc[-2] += hc[-1]; hg, hc = hg[:-1], hc[:-1]
            keep = (hg + hc) > 0
            hg, hc = hg[keep], hc[keep]
            if len(hg) > 1:
                pvals[("count", a, names[k])] = stats.chi2_contingency(np.stack([hg, hc]))[1]
            else:
                assert np.array_equal(np.unique(gc), np.unique(cc)), (a, k)
    # noisy values of the first row of a list: one observer of either team, tick 0
    for a in (0, 2):
        for name, cols in (("robot", (0, 1, 2)), ("post", (0, 3)), ("cross", (0, 3)), ("fcross", (0, 3, 6, 7))):
            k = names.index(name)
            gm, cm = G["counts"][:, 0, a, k] > 0, C["counts"][:, 0, a, k] > 0
            for col in cols:
                gv, cv = G[name][:, 0, a, 0, col][gm], C[name][:, 0, a, 0, col][cm]
                assert len(gv) > 1000 and len(cv) > 300, (name, col, len(gv), len(cv))
                pvals[("value", a, name, col)] = stats.ks_2samp(gv, cv).pvalue
    alpha = 0.01 / len(pvals)
    worst = sorted(pvals.items(), key=lambda kv: kv[1])[:5]
    print("noise statistics: %d tests, Bonferroni threshold %.1e, smallest p-values %s" % (len(pvals), alpha, worst))
    assert all(p > alpha for p in pvals.values()), worst
