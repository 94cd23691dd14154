"""Shared helpers of the parity tests."""
import glob
import os

import numpy as np

from dynenv_b200 import layout as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
OBS_KEYS = ["ball", "robot", "post", "cross", "fcross", "line"]
SEEN_KEYS = ["counts", "num_landmarks", "robots_seen", "balls_seen"]
FULL_KEYS = ["full_ball", "full_self", "full_others"]


def golden_names(prefix="robocup_"):
    return sorted(os.path.basename(p)[len(prefix):-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name, prefix="robocup_"):
    """fixture as a dict; state records written with an older (shorter) layout are zero-padded"""
    from oracle.oracle import pad_words
    g = dict(np.load(os.path.join(GOLDEN, "%s%s.npz" % (prefix, name))))
    for k in ("state0", "states", "state_reset"):
        if k in g and prefix == "robocup_":
            g[k] = pad_words(g[k])
    return g


def golden_cfg(g):
    c = g["cfg"]
    cfg = dict(nPlayers=int(c[0]), obsType=int(c[1]), noiseType=int(c[2]), noiseMag=float(c[3]),
               canFall=bool(c[4]), allowHeadTurn=bool(c[5]), randomInit=bool(c[6]))
    cfg["flags"] = ((L.DYNENV_FLAG_CAN_FALL if cfg["canFall"] else 0) | L.DYNENV_FLAG_USE_OBS_REWARDS
                    | (L.DYNENV_FLAG_RANDOM_INIT if cfg["randomInit"] else 0))
    return cfg


def is_deterministic(cfg):
    """No random draw influences the result: noise off (REALISTIC, magnitude 0), canFall off."""
    return cfg["noiseMag"] == 0 and cfg["noiseType"] == 1 and not cfg["canFall"]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) / (1.0 + np.abs(b))).max()) if a.size else 0.0


def field_errors(wa, wb, tol):
    """names of state fields whose relative error exceeds tol"""
    out = []
    for name, (off, cnt, kind) in L.RC_FIELDS.items():
        e = rel_err(wa[..., off:off + cnt], wb[..., off:off + cnt])
        if e > tol:
            out.append((name, e))
    return out


# ---------------------------------------------------------------- GPU <-> oracle comparison
# Warm-start accumulators of the joint / contact solver (impulses, ~1e5 in a scrum) are compared in the
# unit they act in: impulse x inverse mass (feet: 1/4000) resp. inverse inertia (feet: 1/883333).
# Field lines are clipped to the field of view; a piece that just enters it can be a fraction of a pixel
# long although its end points lie ~700 px from the observer, so its direction amplifies the last bit of
# the float32 head angle (2.4e-7 rad x 700 px / piece length).  The reference then adds +-1 px of noise
# to those very end points.  Lines are therefore compared at 5e-3; every other list at `tol`.
LINE_TOL = 5e-3
IMPULSE_SCALE = {"pivJx": 2.5e-4, "pivJy": 2.5e-4, "rotJ": 1.0 / 883333.3, "arbJn0": 2.5e-4, "arbJt0": 2.5e-4,
                 "arbJn1": 2.5e-4, "arbJt1": 2.5e-4}
def oracle_step_batch(cfg, recs, actions, seed, env_id0=0):
    """Runs the CPU oracle (Philox mode) one step from each state record.  recs: uint32 [N, W]
    (the C-ABI record, float32 words), actions [N, P, 4].  Returns (list of output dicts, records after)."""
    from oracle.oracle import RcOracle, f32_record_to_words, words_to_f32_record
    outs, after = [], []
    for i in range(len(recs)):
        orc = RcOracle(cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"], cfg["allowHeadTurn"],
                       cfg["flags"], seed=seed, env_id=env_id0 + i)
        orc.set_state(f32_record_to_words(recs[i]))
        outs.append(orc.step(actions[i]))
        after.append(orc.get_state())
        assert not orc.arb_overflow()
    return outs, np.array(after)


def compare_step(cfg, sim_out, gpu_state_words, outs, oracle_words, tol=1e-4, line_tol=LINE_TOL):
    """Compares one batched GPU step with per-env oracle results.  Returns a list of
    (env, what, detail[, where]) mismatches: integer / flag / count fields must be equal, floats within
    tol * (1 + |ref|); for the observation lists `where` names the (tick, observer) pairs that differ."""
    P = 2 * cfg["nPlayers"]
    bad = []
    int_mask = np.array(L.rc_int_mask())
    full = cfg["obsType"] == 0
    g = {}
    for k in (FULL_KEYS if full else OBS_KEYS + SEEN_KEYS) + ["rewards", "done", "state_robots", "state_ball",
                                                              "recon_ball", "recon_self", "recon_others", "diag"]:
        g[k] = getattr(sim_out, k).detach().cpu().numpy()
    for e, o in enumerate(outs):
        if g["diag"][e] != 0:
            bad.append((e, "diag", int(g["diag"][e])))
        # state record
        ws, wo = gpu_state_words[e], oracle_words[e]
        for name, (off, cnt, kind) in L.RC_FIELDS.items():
            a, b = ws[off:off + cnt], wo[off:off + cnt]
            if name in IMPULSE_SCALE:  # solver accumulators are compared by the velocity they stand for
                a, b = a * IMPULSE_SCALE[name], b * IMPULSE_SCALE[name]
            if kind == "i":
                if not np.array_equal(a, b):
                    bad.append((e, "state." + name, (a.tolist(), b.tolist())))
            elif rel_err(a, b) > tol:
                bad.append((e, "state." + name, rel_err(a, b)))
        if rel_err(g["rewards"][e], o["rewards"][:P]) > tol:
            bad.append((e, "rewards", rel_err(g["rewards"][e], o["rewards"][:P])))
        if bool(g["done"][e]) != bool(o["done"]):
            bad.append((e, "done", None))
        for k, sl in (("state_robots", np.s_[:P]), ("state_ball", np.s_[:]), ("recon_ball", np.s_[:P]),
                      ("recon_self", np.s_[:P]), ("recon_others", np.s_[:P, :P - 1])):
            ref = o[k][sl]
            got = g[k][e][sl] if k == "recon_others" else g[k][e]
            if rel_err(got, ref) > tol:
                bad.append((e, k, rel_err(got, ref)))
        if full:
            for k in FULL_KEYS:
                ref = o[k][:, :P]
                got = g[k][e]
                if k == "full_others":
                    ref, got = ref[:, :, :P - 1], got[:, :, :P - 1]
                if rel_err(got, ref) > tol:
                    bad.append((e, k, rel_err(got, ref)))
        else:
            for k in SEEN_KEYS:
                ref = o[k][:, :P]
                got = g[k][e]
                if k == "robots_seen":
                    ref, got = ref[..., :P - 1], got[..., :P - 1]
                if not np.array_equal(ref, got):
                    bad.append((e, k, np.argwhere(ref != got).tolist()))  # [tick, observer(, type / slot)] of every difference
            for k in OBS_KEYS:
                ref = o[k][:, :P]
                got = g[k][e]
                lim = max(tol, line_tol if k == "line" else 0)
                if rel_err(got, ref) > lim:
                    where = [(t, a) for t in range(ref.shape[0]) for a in range(P) if rel_err(got[t, a], ref[t, a]) > lim]
                    bad.append((e, k, rel_err(got, ref), where))
    return bad


# ---------------------------------------------------------------- certificates: a disagreement must be a PROVEN tie
# The kernels compute in float32, the oracle in float64; both apply the reference's hard thresholds.  A state that sits
# ON a threshold is decided by the last bit on either side (and in the reference itself).  The parity tests accept a
# disagreement only with a certificate that it is such a tie, computed from the oracle's own state:
#
#  on_line    the observer stands within ON_LINE_EPS px of the supporting line of a field line (every penalised robot
#             does: the penalty spots lie ON the side lines, RoboCupEnvironment.py:139-151): the line passes through the
#             apex of its field of view, isLineInArea's edge tests (cutils.py:751-821) are cross products of ~0.  Covers
#             that observer's line rows, its line count / numLandMarks, and its observation reward within the bound of
#             processSeens (:1563-1575): 0.0025 per landmark, averaged over the 5 ticks.
#  argmin_tie two robots of one team are equidistant from the ball within ARGMIN_EPS (relative) when the step starts
#             (the ten-fold symmetric scrum): closestID (:722-732) is a tie.  Covers the approach / ball rewards of that
#             environment (:1003-1007, :707-716) and the `closest` flags.
#  stiff      hand-made deep inter-penetration (robots placed into each other): impulses of ~1e7 amplify float32
#             rounding through ten Gauss-Seidel sweeps; continuous state is compared at STIFF_TOL there, everything
#             discrete stays exact.
ON_LINE_EPS = 1e-3
ARGMIN_EPS = 1e-9
STIFF_TOL = 5e-3
_LINE_X = (70.0, 970.0, 520.0, 130.0, 910.0)
_LINE_Y = (70.0, 670.0, 260.0, 480.0)
REWARD_FIELDS = ("rewards", "state.episodeRewards", "state.episodePosRewards", "state.episodeObsRewards")


def robot_positions(words, P):
    """[P, 2] Robot.getPos of a state record (float64 words)"""
    px = words[L.RC_S_PX:L.RC_S_PX + 2 * P] + words[L.RC_S_PXL:L.RC_S_PXL + 2 * P]
    py = words[L.RC_S_PY:L.RC_S_PY + 2 * P] + words[L.RC_S_PYL:L.RC_S_PYL + 2 * P]
    return np.stack([(px[0::2] + px[1::2]) / 2, (py[0::2] + py[1::2]) / 2], -1)


def on_line_observers(words_before, words_after, P):
    """robots that stand ON a field line's supporting line when the step ends and either stood there all step long or
    are penalised (teleported onto their penalty spot during the step: every tick from then on sees the line through
    the apex; the ticks before cannot disagree on it)"""
    out = set()
    pb, pa = robot_positions(words_before, P), robot_positions(words_after, P)
    for a in range(P):
        d = min(min(abs(pa[a, 0] - x) for x in _LINE_X), min(abs(pa[a, 1] - y) for y in _LINE_Y))
        penalised = int(words_after[L.RC_S_RFLAGS + a]) & L.RC_RF_PENALIZED
        if d < ON_LINE_EPS and (np.abs(pa[a] - pb[a]).max() < ON_LINE_EPS or penalised):
            out.add(a)
    return out


def argmin_tie(words_before, P):
    """True if the closest-robot argmin of a team is a tie in the step's initial state"""
    pos = robot_positions(words_before, P)
    ball = np.array([words_before[L.RC_S_PX + 20] + words_before[L.RC_S_PXL + 20], words_before[L.RC_S_PY + 20] + words_before[L.RC_S_PYL + 20]])
    for team in (pos[:P // 2], pos[P // 2:]):
        d2 = np.sort(((team - ball) ** 2).sum(1))
        if len(d2) > 1 and d2[1] - d2[0] <= ARGMIN_EPS * (1.0 + d2[1]):
            return True
    return False


def certify(bad, words_before, words_after, P, outs=None, gpu_out=None):
    """Splits the mismatches of compare_step into (explained, unexplained).  words_before / words_after: the oracle's
    state records [n, W] (float64 words) around the step.  explained entries get the certificate's name appended."""
    explained, unexplained = [], []
    by_env = {}
    for b in bad:
        by_env.setdefault(b[0], []).append(b)
    for e, items in by_env.items():
        online = on_line_observers(words_before[e], words_after[e], P)
        tie = argmin_tie(words_before[e], P)
        line_obs = set()  # observers with a certified line disagreement
        for b in items:
            if b[1] == "line" and all(a in online for (t, a) in b[3]):
                line_obs |= set(a for (t, a) in b[3])
        for b in items:
            k = b[1]
            cert = None
            if k == "line" and all(a in online for (t, a) in b[3]):
                cert = "on_line"
            elif k == "counts" and all(w[1] in online and w[2] == 5 for w in b[2]):
                cert = "on_line"
            elif k == "num_landmarks" and all(w[1] in online for w in b[2]):
                cert = "on_line"
            elif k in REWARD_FIELDS and line_obs:
                cert = "on_line"  # observation reward of the on-line observers (bounded below)
            elif tie and (k in REWARD_FIELDS or k in ("recon_ball", "state.closestID", "ball", "full_ball")):
                cert = "argmin_tie"
            (explained if cert else unexplained).append(b + (cert,) if cert else b)
        # the reward differences an on-line certificate covers are bounded: 0.0025 per landmark and tick / 5 ticks
        if line_obs and outs is not None and gpu_out is not None:
            ref = np.asarray(outs[e]["rewards"][:P], np.float64)
            got = gpu_out.rewards.detach().cpu().numpy()[e][:P].astype(np.float64)
            for a in range(P):
                lim = (0.0025 * 11 if a in line_obs else 0.0) + 1e-4 * (1 + abs(ref[a]))
                if abs(got[a] - ref[a]) > lim and not tie:
                    unexplained.append((e, "rewards", float(abs(got[a] - ref[a])), "beyond the observation-reward bound"))
    return explained, unexplained


def mismatch_table(rows):
    """text table: scenario, environments, mismatching environments, per certificate, unexplained"""
    lines = ["%-34s %5s %9s %8s %11s %6s %12s" % ("scenario", "envs", "mismatch", "on_line", "argmin_tie", "stiff", "unexplained")]
    for name, n, n_bad, certs, n_un in rows:
        lines.append("%-34s %5d %9d %8d %11d %6d %12d" % (name, n, n_bad, certs.get("on_line", 0), certs.get("argmin_tie", 0),
                                                          certs.get("stiff", 0), n_un))
    return "\n".join(lines)


# ---------------------------------------------------------------- Driving: GPU <-> oracle comparison
DRV_OBS_KEYS = ["self", "car", "obs", "ped", "lane"]
# warm-start impulses are compared by the velocity they stand for (lightest car: 1200)
DRV_IMPULSE_SCALE = {"arbJn0": 1.0 / 1200, "arbJt0": 1.0 / 1200, "arbJn1": 1.0 / 1200, "arbJt1": 1.0 / 1200}


def drv_golden_cfg(g):
    c = g["cfg"]
    return dict(nPlayers=int(c[0]), obsType=int(c[1]), noiseType=int(c[2]), noiseMag=float(c[3]), continuous=bool(c[4]))


def drv_oracle_step_batch(cfg, recs, actions, seed, env_id0=0):
    """CPU oracle (Philox mode), one step from each Driving state record (uint32 [N, W]); actions [N, nCars, 2]."""
    from oracle.oracle import DrvOracle, drv_f32_record_to_words
    outs, after = [], []
    for i in range(len(recs)):
        orc = DrvOracle(cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"], cfg["continuous"], seed=seed,
                        env_id=env_id0 + i)
        orc.set_state(drv_f32_record_to_words(recs[i]))
        outs.append(orc.step(actions[i]))
        after.append(orc.get_state())
        assert not orc.arb_overflow()
    return outs, np.array(after)


def drv_compare_obs(e, g, o, N, tol, bad):
    if not np.array_equal(g["counts"][e][:N], o["counts"][:N]):
        bad.append((e, "counts", (g["counts"][e][:N].tolist(), o["counts"][:N].tolist())))
        return
    for k in DRV_OBS_KEYS:
        err = rel_err(g[k][e][:N], o[k][:N])
        if err > tol:
            bad.append((e, k, err))


def drv_compare_step(cfg, sim_out, gpu_state_words, outs, oracle_words, tol=1e-4):
    """One batched GPU Driving step against per-env oracle results: integer / flag / count fields equal, floats
    within tol * (1 + |ref|).  Returns a list of (env, what, detail) mismatches."""
    N = min(cfg["nPlayers"], L.DRV_MAXCARS)
    bad = []
    g = {"self": sim_out.self_}
    for k in DRV_OBS_KEYS[1:] + ["counts", "rewards", "done", "diag", "state_cars", "state_obs", "state_ped", "n_obs", "n_ped",
                                 "episode_r", "episode_p_r", "episode_g"]:
        g[k] = getattr(sim_out, k)
    g = {k: v.detach().cpu().numpy() for k, v in g.items()}
    for e, o in enumerate(outs):
        if g["diag"][e] != 0:
            bad.append((e, "diag", int(g["diag"][e])))
        ws, wo = gpu_state_words[e], oracle_words[e]
        for name, (off, cnt, kind) in L.DRV_FIELDS.items():
            a, b = ws[off:off + cnt], wo[off:off + cnt]
            if name in DRV_IMPULSE_SCALE:
                a, b = a * DRV_IMPULSE_SCALE[name], b * DRV_IMPULSE_SCALE[name]
            if kind == "i":
                if not np.array_equal(a, b):
                    bad.append((e, "state." + name, (a.tolist(), b.tolist())))
            elif rel_err(a, b) > tol:
                bad.append((e, "state." + name, rel_err(a, b)))
        if rel_err(g["rewards"][e], o["rewards"][:N]) > tol:
            bad.append((e, "rewards", rel_err(g["rewards"][e], o["rewards"][:N])))
        if bool(g["done"][e]) != bool(o["done"]):
            bad.append((e, "done", None))
        if o["done"]:
            for k in ("episode_r", "episode_p_r"):
                if rel_err(g[k][e], o[k][:N]) > tol * 10:
                    bad.append((e, k, rel_err(g[k][e], o[k][:N])))
            if not np.array_equal(g["episode_g"][e], o["episode_g"]):
                bad.append((e, "episode_g", None))
        for k, ref in (("state_cars", o["state_cars"][:N]), ("state_obs", o["state_obs"]), ("state_ped", o["state_ped"])):
            if rel_err(g[k][e], ref) > tol:
                bad.append((e, k, rel_err(g[k][e], ref)))
        if int(g["n_obs"][e]) != int(o["nobs"]) or int(g["n_ped"][e]) != int(o["nped"]):
            bad.append((e, "n_obs/n_ped", None))
        drv_compare_obs(e, g, o, N, tol, bad)
    return bad
