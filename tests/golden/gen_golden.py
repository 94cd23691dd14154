"""Generates tests/golden/robocup_*.npz by running the REFERENCE's own Python
(/root/reference/DynEnv, unmodified) on the restated physics shim (oracle/shim).

Run here only (the GPU box has no /root/reference):  python tests/golden/gen_golden.py

Each fixture holds, for one scenario: the configuration, the initial state record (float64
words in the C-ABI layout), the per-step actions, the uniforms the reference consumed per
step (RNG tape), and what the reference produced per step: state record, rewards, done,
'Full State', and (for the steps listed in `obs_steps`) the padded observations.
Random-play scenarios start from the reference's own reset(); contact scenarios move
bodies of a reset scene (ball at a foot, robots face to face, robot at a post, ball flying
at a post / over the lines, defenders in the box) to exercise every collision callback
and rule.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dynenv_b200 import layout as L  # noqa: E402
from oracle import ref_state  # noqa: E402
from oracle.ref_loader import UniformTape, load_reference  # noqa: E402

F = L.RC_FIELDS
OUT = os.path.dirname(os.path.abspath(__file__))
CAPS = dict(ball=(L.RC_CAP_BALL, L.RC_F_BALL), robot=(L.RC_CAP_ROBOT, L.RC_F_ROBOT),
            post=(L.RC_CAP_POST, L.RC_F_POST), cross=(L.RC_CAP_CROSS, L.RC_F_CROSS),
            fcross=(L.RC_CAP_FCROSS, L.RC_F_FCROSS), line=(L.RC_CAP_LINE, L.RC_F_LINE))
TYPES = ["ball", "robot", "post", "cross", "fcross", "line"]


def ref_obs_to_arrays(obs, P, full):
    """reference nested observation (list[T] of list[P] of tuples) -> padded arrays"""
    T = len(obs)
    out = {}
    if full:
        out["full_ball"] = np.zeros((T, 10, 4), np.float32)
        out["full_self"] = np.zeros((T, 10, 8), np.float32)
        out["full_others"] = np.zeros((T, 10, 9, 6), np.float32)
        for t in range(T):
            for r in range(P):
                (b, o), (s,), _ = obs[t][r]
                out["full_ball"][t, r] = b[0]
                out["full_self"][t, r] = s[0]
                out["full_others"][t, r, :P - 1] = o
        return out
    for k, (c, f) in CAPS.items():
        out[k] = np.zeros((T, 10, c, f), np.float32)
    out["counts"] = np.zeros((T, 10, 6), np.int32)
    out["num_landmarks"] = np.zeros((T, 10), np.int32)
    out["robots_seen"] = np.zeros((T, 10, 9), np.int32)
    out["balls_seen"] = np.zeros((T, 10), np.int32)
    for t in range(T):
        for r in range(P):
            (b, ro), (po, cr, fc, li), (nl, rs, bs) = obs[t][r]
            for i, (k, a) in enumerate(zip(TYPES, [b, ro, po, cr, fc, li])):
                n = min(a.shape[0] if a.ndim == 2 else 0, CAPS[k][0])
                out["counts"][t, r, i] = n
                if n:
                    out[k][t, r, :n] = a[:n]
            out["num_landmarks"][t, r] = nl
            out["robots_seen"][t, r, :P - 1] = rs
            out["balls_seen"][t, r] = bs
    return out


def place_robot(w, r, x, y, ang):
    for f in (0, 1):
        i = 2 * r + f
        w[F["px"][0] + i] = x
        w[F["py"][0] + i] = y
        w[F["angle"][0] + i] = ang
        ly = 10.0 if f == 0 else -10.0
        c, s = np.cos(ang), np.sin(ang)
        w[F["cax"][0] + i] = x + (-10.0) * c - ly * s
        w[F["cay"][0] + i] = y + (-10.0) * s + ly * c
        w[F["cbx"][0] + i] = x + 10.0 * c - ly * s
        w[F["cby"][0] + i] = y + 10.0 * s + ly * c
    w[F["prevX"][0] + r] = x
    w[F["prevY"][0] + r] = y


def place_ball(w, x, y, vx=0.0, vy=0.0, spin=0.0):
    w[F["px"][0] + 20] = x
    w[F["py"][0] + 20] = y
    w[F["vx"][0] + 20] = vx
    w[F["vy"][0] + 20] = vy
    w[F["w"][0] + 20] = spin
    for n, v in (("cax", x), ("cay", y), ("cbx", x), ("cby", y)):
        w[F[n][0] + 20] = v
    w[F["ballPrevX"][0]] = x
    w[F["ballPrevY"][0]] = y


# scenario -> (config, craft(words, P) or None, action script(step, P, rng) -> [P,4], steps)
def _rand_actions(step, P, rng, head_cont=False):
    a = np.stack([rng.integers(0, 5, P), rng.integers(0, 3, P), rng.integers(0, 3, P),
                  rng.integers(0, 7, P)], 1).astype(np.float64)
    if head_cont:
        a[:, 3] = rng.uniform(-3, 3, P)
    return a


def _const_actions(rows):
    def f(step, P, rng):
        a = np.zeros((P, 4))
        a[:, 3] = 3
        for r, row in rows.items():
            a[r] = row
        return a
    return f


def craft_ball_at_foot(w, P):
    place_robot(w, 0, 460.0, 372.0, 0.0)
    place_ball(w, 495.0, 370.0)


def craft_kick(w, P):
    place_robot(w, 0, 470.0, 362.0, 0.1)
    place_ball(w, 498.0, 352.0)


def craft_face_to_face(w, P):
    n = P // 2
    place_robot(w, 0, 400.0, 300.0, 0.0)
    place_robot(w, n, 446.0, 303.0, np.pi)
    if n > 1:
        place_robot(w, 1, 400.0, 500.0, 0.3)
        place_robot(w, n + 1, 440.0, 520.0, np.pi + 0.2)


def craft_robot_at_post(w, P):
    place_robot(w, 0, 100.0, 447.0, np.pi)
    n = P // 2
    place_robot(w, n, 940.0, 292.0, 0.0)


def craft_ball_to_post(w, P):
    place_ball(w, 110.0, 445.0, vx=-300.0, vy=20.0, spin=3.0)


def craft_goal(w, P):
    place_ball(w, 100.0, 380.0, vx=-400.0, vy=5.0)
    w[F["lastKicked"][0]] = P // 2
    w[F["nLastKicked"][0]] = 1


def craft_throw_in(w, P):
    place_ball(w, 600.0, 90.0, vx=30.0, vy=-350.0)
    w[F["lastKicked"][0]] = 0
    w[F["nLastKicked"][0]] = 1


def craft_corner(w, P):
    place_ball(w, 960.0, 600.0, vx=350.0, vy=30.0)
    w[F["lastKicked"][0]] = P // 2
    w[F["nLastKicked"][0]] = 1


def craft_defenders(w, P):
    n = P // 2
    place_robot(w, 0, 90.0, 300.0, 0.0)
    place_robot(w, 1, 90.0, 370.0, 0.0)
    place_robot(w, 2, 150.0, 430.0, np.pi)
    place_robot(w, n, 1030.0, 100.0, 0.0)  # about to leave the field


def craft_scrum(w, P):
    n = P // 2
    for k in range(n):
        ang = 2 * np.pi * k / P
        place_robot(w, k, 520.0 + 55.0 * np.cos(ang), 370.0 + 55.0 * np.sin(ang), ang + np.pi)
        ang2 = 2 * np.pi * (k + n) / P
        place_robot(w, n + k, 520.0 + 55.0 * np.cos(ang2), 370.0 + 55.0 * np.sin(ang2), ang2 + np.pi)


def _cfg(nPlayers, obs=1, noise=1, mag=0.0, canFall=False, head=False, randomInit=False):
    return dict(nPlayers=nPlayers, obsType=obs, noiseType=noise, noiseMag=mag, canFall=canFall,
                allowHeadTurn=head, randomInit=randomInit)


SCENARIOS = {
    # name: (config, craft, actions, steps, seed)
    "c1_1v1_full": (_cfg(1, obs=0), None, _rand_actions, 40, 42),
    "rand_5v5_clean": (_cfg(5), None, _rand_actions, 30, 7),
    "rand_5v5_noisy": (_cfg(5, mag=2.0, canFall=True), None, _rand_actions, 30, 8),
    "rand_5v5_randomnoise": (_cfg(5, noise=0, mag=3.0, canFall=True, randomInit=True), None, _rand_actions, 30, 9),
    "rand_3v3_headturn": (_cfg(3, mag=5.0, canFall=True, head=True), None,
                          lambda s, P, g: _rand_actions(s, P, g, True), 30, 10),
    "ball_at_foot": (_cfg(2), craft_ball_at_foot, _const_actions({0: [3, 0, 0, 3]}), 12, 11),
    "kick": (_cfg(2), craft_kick, _const_actions({0: [0, 0, 2, 3]}), 8, 12),
    "face_to_face": (_cfg(2), craft_face_to_face,
                     _const_actions({0: [3, 0, 0, 3], 2: [3, 0, 0, 3], 1: [3, 0, 0, 4], 3: [3, 0, 0, 2]}), 14, 13),
    "face_to_face_fall": (_cfg(2, mag=2.0, canFall=True), craft_face_to_face,
                          _const_actions({0: [3, 0, 0, 3], 2: [3, 0, 0, 3], 1: [3, 0, 0, 4], 3: [3, 0, 0, 2]}), 40, 14),
    "robot_at_post": (_cfg(1), craft_robot_at_post, _const_actions({0: [3, 0, 0, 3], 1: [3, 0, 0, 3]}), 12, 15),
    "robot_at_post_fall": (_cfg(1, mag=1.0, canFall=True), craft_robot_at_post,
                           _const_actions({0: [3, 0, 0, 3], 1: [3, 0, 0, 3]}), 60, 16),
    "ball_to_post": (_cfg(1), craft_ball_to_post, _const_actions({}), 6, 17),
    "goal": (_cfg(2), craft_goal, _const_actions({}), 4, 18),
    "throw_in": (_cfg(2), craft_throw_in, _const_actions({}), 4, 19),
    "corner": (_cfg(2), craft_corner, _const_actions({}), 4, 20),
    "defenders": (_cfg(5), craft_defenders, _const_actions({5: [3, 0, 0, 3], 2: [3, 0, 0, 3]}), 8, 21),
    "scrum": (_cfg(5), craft_scrum, _const_actions({k: [3, 0, 0, 3] for k in range(10)}), 25, 22),
    "scrum_fall": (_cfg(5, mag=2.0, canFall=True), craft_scrum,
                   _const_actions({k: [3, 0, 0, 3] for k in range(10)}), 40, 23),
}


def make_ref_env(ns, cfg, tape):
    ns.RoboCupEnvironment.canFall = cfg["canFall"]
    ns.RoboCupEnvironment.randomInit = cfg["randomInit"]
    with tape.installed():
        env = ns.RoboCupEnvironment(cfg["nPlayers"], observationType=ns.ObservationType(cfg["obsType"]),
                                    noiseType=ns.NoiseType(cfg["noiseType"]), noiseMagnitude=cfg["noiseMag"],
                                    allowHeadTurn=cfg["allowHeadTurn"])
    tape.take_log()
    return env


def build_record(ns, cfg, craft, actions_fn, steps, seed):
    """-> (the fixture's arrays, statistics of the run); also what tests/test_oracle_live.py checks the oracle against on
    scenarios that are not among the committed ones"""
    tape = UniformTape(seed)
    env = make_ref_env(ns, cfg, tape)
    P = 2 * min(cfg["nPlayers"], 5)
    full = cfg["obsType"] == 0
    with tape.installed():
        obs0 = env.reset()
    reset_tape = tape.take_log()
    # the tape is the PCG64(seed) stream of UniformTape: store where each piece starts, not the
    # uniforms themselves (oracle.ref_loader.UniformTape(seed) regenerates them)
    n_ctor = tape.count - len(reset_tape)
    rec = {"tape_seed": np.array(seed), "tape_reset_off": np.array([n_ctor, tape.count]),
           "tape_reset_sum": np.array(reset_tape.sum())}
    state_reset = ref_state.robocup_get_state(env)
    for k, v in ref_obs_to_arrays(obs0, P, full).items():
        rec["reset_" + k] = v
    rec["state_reset"] = state_reset
    w0 = state_reset.copy()
    if craft is not None:
        craft(w0, P)
        ref_state.robocup_set_state(env, w0, ns)
        assert np.array_equal(ref_state.robocup_get_state(env), w0)
    rec["state0"] = w0
    rng = np.random.default_rng(seed + 100)
    acts, tapes, tape_off, states, rewards, dones, fs_r, fs_b = [], [], [tape.count], [], [], [], [], []
    obs_steps = sorted(set([0, 1, steps // 2, steps - 1]))
    stats = dict(maxarb=0, fallen=0, penal=0, goals=0)
    for step in range(steps):
        a = actions_fn(step, P, rng)
        with tape.installed():
            obs, rew, fin, info = env.step(a.copy())
        lg = tape.take_log()
        acts.append(a)
        tapes.append(lg)
        tape_off.append(tape.count)
        w = ref_state.robocup_get_state(env)
        states.append(w)
        rw = np.zeros(10)
        rw[:P] = rew
        rewards.append(rw)
        dones.append(bool(fin))
        r6 = np.zeros((10, 6), np.float32)
        r6[:P] = info["Full State"][0]
        fs_r.append(r6)
        fs_b.append(info["Full State"][1])
        if step in obs_steps:
            for k, v in ref_obs_to_arrays(obs, P, full).items():
                rec["obs%d_%s" % (step, k)] = v
        stats["maxarb"] = max(stats["maxarb"], int(w[F["nArb"][0]]))
        fl = w[F["rflags"][0]:F["rflags"][0] + 10].astype(int)
        stats["fallen"] += int(((fl & 4) > 0).sum())
        stats["penal"] += int(((fl & 8) > 0).sum())
        stats["goals"] = int(w[F["goals"][0]] + w[F["goals"][0] + 1])
    rec.update(actions=np.array(acts), tape_sum=np.array(sum(t.sum() for t in tapes)),
               tape_off=np.array(tape_off), states=np.array(states), rewards=np.array(rewards),
               dones=np.array(dones), fs_robots=np.array(fs_r), fs_ball=np.array(fs_b),
               obs_steps=np.array(obs_steps),
               cfg=np.array([cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"],
                             int(cfg["canFall"]), int(cfg["allowHeadTurn"]), int(cfg["randomInit"])],
                            dtype=np.float64))
    return rec, stats


def run_scenario(ns, name):
    cfg, craft, actions_fn, steps, seed = SCENARIOS[name]
    rec, stats = build_record(ns, cfg, craft, actions_fn, steps, seed)
    np.savez_compressed(os.path.join(OUT, "robocup_%s.npz" % name), **rec)
    print("%-24s steps=%3d max cached pairs=%2d fallen-robot-steps=%3d penalized-robot-steps=%3d goals=%d"
          % (name, steps, stats["maxarb"], stats["fallen"], stats["penal"], stats["goals"]))


if __name__ == "__main__":
    ns = load_reference()
    names = sys.argv[1:] or list(SCENARIOS)
    for n in names:
        run_scenario(ns, n)
