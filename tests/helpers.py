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


def compare_step(cfg, sim_out, gpu_state_words, outs, oracle_words, tol=1e-4):
    """Compares one batched GPU step with per-env oracle results.  Returns a list of
    (env, what, detail) mismatches: integer / flag / count fields must be equal, floats within
    tol * (1 + |ref|)."""
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
                    bad.append((e, k, np.argwhere(ref != got)[:4].tolist()))
            for k in OBS_KEYS:
                ref = o[k][:, :P]
                got = g[k][e]
                if rel_err(got, ref) > max(tol, LINE_TOL if k == "line" else 0):
                    bad.append((e, k, rel_err(got, ref)))
    return bad


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
