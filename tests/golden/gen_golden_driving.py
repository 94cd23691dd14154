"""Generates tests/golden/driving_*.npz by running the REFERENCE's own Python (/root/reference/DynEnv,
unmodified) on the restated physics shim (oracle/shim).  Run here only: python tests/golden/gen_golden_driving.py

Same structure as gen_golden.py: configuration, initial state record (float64 words, C-ABI layout), actions, the
uniforms the reference consumed (as offsets into the PCG64 tape), and per step: state record, rewards, done,
'Full State', observations.  The `continuous` scenario patches processAction with acc = action[0],
steer = action[1], because the reference's own continuous branch raises UnboundLocalError
(DrivingEnvironment.py:360-368)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dynenv_b200 import layout as L  # noqa: E402
from oracle import ref_state  # noqa: E402
from oracle.ref_loader import UniformTape, load_reference  # noqa: E402

F = L.DRV_FIELDS
OUT = os.path.dirname(os.path.abspath(__file__))
NC = L.DRV_MAXCARS


def ref_obs_to_arrays(obs_tick, N, full):
    """one tick of the reference's observation (list[N] of tuples) -> padded arrays"""
    out = {"self": np.zeros((NC, L.DRV_F_SELF), np.float32), "car": np.zeros((NC, L.DRV_CAP_CAR, L.DRV_F_CAR), np.float32),
           "obs": np.zeros((NC, L.DRV_CAP_OBS, L.DRV_F_OBS), np.float32), "ped": np.zeros((NC, L.DRV_CAP_PED, L.DRV_F_PED), np.float32),
           "lane": np.zeros((NC, L.DRV_CAP_LANE, L.DRV_F_LANE_FULL), np.float32), "counts": np.zeros((NC, 4), np.int32)}
    for a in range(N):
        (cars, obst, peds), (self_, lanes), _ = obs_tick[a]
        out["self"][a] = self_[0]
        for i, (k, arr) in enumerate((("car", cars), ("obs", obst), ("ped", peds), ("lane", lanes))):
            n = arr.shape[0] if arr.ndim == 2 else 0
            n = min(n, out[k].shape[1])
            out["counts"][a, i] = n
            if n:
                out[k][a, :n, :arr.shape[1]] = arr[:n]
    return out


def _rand_actions(step, N, rng, continuous=False):
    if continuous:
        return rng.uniform(-3, 3, (N, 2))
    return rng.integers(0, 3, (N, 2)).astype(np.float64)


def _const(rows, default=(1, 1)):
    def f(step, N, rng):
        a = np.tile(np.array(default, dtype=np.float64), (N, 1))
        for r, row in rows.items():
            a[r] = row
        return a
    return f


def place_car(w, i, x, y, ang, vx=0.0, vy=0.0):
    for n, v in (("cpx", x), ("cpy", y), ("cang", ang), ("cvx", vx), ("cvy", vy), ("cw", 0.0), ("cdirx", np.cos(ang)),
                 ("cdiry", np.sin(ang)), ("cprevx", x), ("cprevy", y)):
        w[F[n][0] + i] = v
    w[F["cflags"][0] + i] = 0


def place_ped(w, i, x, y, vx=0.0, vy=0.0):
    for n, v in (("ppx", x), ("ppy", y), ("pvx", vx), ("pvy", vy)):
        w[F[n][0] + i] = v


def craft_car_car(w, N):
    place_car(w, 0, 857.5, 300.0, np.pi / 2, 0.0, 60.0)
    place_car(w, 1, 859.0, 345.0, -np.pi / 2, 0.0, -60.0)
    if N > 3:  # two cars in their right lanes, one rear-ending the other
        place_car(w, 2, 892.5, 700.0, -np.pi / 2, 0.0, -80.0)
        place_car(w, 3, 893.5, 660.0, -np.pi / 2, 0.0, -5.0)


def craft_car_ped(w, N):
    place_car(w, 0, 857.5, 380.0, np.pi / 2, 0.0, 50.0)
    place_ped(w, 0, 858.0, 402.0)
    place_car(w, 1, 892.5, 600.0, -np.pi / 2, 0.0, -0.5)   # too slow: the pair is ignored
    place_ped(w, 1, 893.0, 580.0)
    place_car(w, 2, 400.0, 517.5, 0.0, 40.0, 0.0)          # pedestrian hit from the side, car not "towards" it
    place_ped(w, 2, 415.0, 532.0, 0.0, -6.0)


def craft_car_static(w, N):
    place_car(w, 0, 785.0, 300.0, np.pi, -50.0, 0.0)        # into the left building
    ox, oy = w[F["opx"][0]], w[F["opy"][0]]
    place_car(w, 1, ox - 28.0, oy + 2.0, 0.0, 45.0, 0.0)    # into obstacle 0


def craft_goal_offroad(w, N):
    for i in range(min(N, 3)):
        gx, gy = w[F["cgoalx"][0] + i], w[F["cgoaly"][0] + i]
        vertical = abs(gx - 875.0) < 1e-9
        if vertical:
            y = gy + (103.0 if gy < 500 else -103.0)
            place_car(w, i, 857.5, y, np.pi / 2 if gy > 500 else -np.pi / 2, 0.0, 40.0 if gy > 500 else -40.0)
        else:
            x = gx + (103.0 if gx < 875 else -103.0)
            place_car(w, i, x, 517.5 if gx > 875 else 482.5, 0.0 if gx > 875 else np.pi, 40.0 if gx > 875 else -40.0, 0.0)
    if N > 3:
        place_car(w, 3, 957.0, 300.0, 0.0, 30.0, 0.0)        # leaves the road -> crash
    if N > 4:
        place_car(w, 4, 1755.0, 482.0, 0.0, 30.0, 0.0)       # out of the field: velocity zeroed


def _cfg(n, obs=1, noise=1, mag=0.0, continuous=False):
    return dict(nPlayers=n, obsType=obs, noiseType=noise, noiseMag=mag, continuous=continuous)


SCENARIOS = {
    "rand_10_noisy": (_cfg(10, mag=2.0), None, _rand_actions, 30, 31),
    "rand_10_clean": (_cfg(10), None, _rand_actions, 30, 32),
    "rand_10_randomnoise": (_cfg(10, noise=0, mag=3.0), None, _rand_actions, 25, 33),
    "rand_5_full": (_cfg(5, obs=0), None, _rand_actions, 30, 34),
    "rand_10_continuous": (_cfg(10, mag=2.0, continuous=True), None, lambda s, N, g: _rand_actions(s, N, g, True), 25, 35),
    "accelerate_10": (_cfg(10, mag=1.0), None, _const({}, (2, 1)), 40, 36),
    "car_car": (_cfg(6), craft_car_car, _const({}), 8, 37),
    "car_car_noisy": (_cfg(6, mag=2.0), craft_car_car, _const({}), 8, 38),
    "car_ped": (_cfg(5), craft_car_ped, _const({}), 10, 39),
    "car_static": (_cfg(4), craft_car_static, _const({}), 8, 40),
    "goal_offroad": (_cfg(6), craft_goal_offroad, _const({}), 8, 41),
}


def make_ref_env(ns, cfg, tape):
    if cfg["continuous"]:
        def processAction(self, action, car):
            acc, steer = action[0], action[1]
            car.accelerate(acc.item(), True)
            if steer != 0:
                car.turn(steer.item())
        ns.DrivingEnvironment.processAction = processAction
    elif hasattr(ns, "_orig_processAction"):
        ns.DrivingEnvironment.processAction = ns._orig_processAction
    with tape.installed():
        env = ns.DrivingEnvironment(cfg["nPlayers"], observationType=ns.ObservationType(cfg["obsType"]),
                                    noiseType=ns.NoiseType(cfg["noiseType"]), noiseMagnitude=cfg["noiseMag"],
                                    continuousActions=cfg["continuous"])
    tape.take_log()
    return env


def build_record(ns, cfg, craft, actions_fn, steps, seed):
    """-> (the fixture's arrays, statistics of the run, uniforms drawn); also what tests/test_oracle_live.py checks the
    oracle against on scenarios that are not among the committed ones"""
    tape = UniformTape(seed)
    env = make_ref_env(ns, cfg, tape)
    N = min(cfg["nPlayers"], NC)
    full = cfg["obsType"] == 0
    with tape.installed():
        obs0 = env.reset()
    reset_tape = tape.take_log()
    n_ctor = tape.count - len(reset_tape)
    rec = {"tape_seed": np.array(seed), "tape_reset_off": np.array([n_ctor, tape.count]), "tape_reset_sum": np.array(reset_tape.sum())}
    state_reset = ref_state.driving_get_state(env)
    for k, v in ref_obs_to_arrays(obs0[0], N, full).items():
        rec["reset_" + k] = v
    rec["state_reset"] = state_reset
    w0 = state_reset.copy()
    if craft is not None:
        craft(w0, N)
        ref_state.driving_set_state(env, w0, ns)
        w0 = ref_state.driving_get_state(env)
    rec["state0"] = w0
    rng = np.random.default_rng(seed + 100)
    acts, tape_off, states, rewards, dones, fs = [], [tape.count], [], [], [], []
    stats = dict(maxarb=0, crashed=0, finished=0, dead=0)
    for step in range(steps):
        a = actions_fn(step, N, rng)
        with tape.installed():
            obs, rew, fin, info = env.step(a.copy())
        tape.take_log()
        acts.append(a)
        tape_off.append(tape.count)
        w = ref_state.driving_get_state(env)
        states.append(w)
        rw = np.zeros(NC)
        rw[:N] = rew
        rewards.append(rw)
        dones.append(bool(fin))
        c7 = np.zeros((NC, 7), np.float32)
        c7[:N] = info["Full State"][0]
        fs.append(c7)
        for k, v in ref_obs_to_arrays(obs[0], N, full).items():
            rec.setdefault("obs_" + k, []).append(v)
        fl = w[F["cflags"][0]:F["cflags"][0] + NC].astype(int)
        stats["maxarb"] = max(stats["maxarb"], int(w[F["nArb"][0]]))
        stats["crashed"] = int(((fl & 2) > 0).sum())
        stats["finished"] = int(((fl & 1) > 0).sum())
        stats["dead"] = int(((w[F["pflags"][0]:F["pflags"][0] + L.DRV_MAXPED].astype(int) & 1) > 0).sum())
    for k in list(rec):
        if k.startswith("obs_"):
            rec[k] = np.array(rec[k])
    rec.update(actions=np.array(acts), tape_off=np.array(tape_off), states=np.array(states), rewards=np.array(rewards),
               dones=np.array(dones), fs_cars=np.array(fs),
               cfg=np.array([cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"], int(cfg["continuous"])], dtype=np.float64))
    return rec, stats, tape.count


def run_scenario(ns, name):
    cfg, craft, actions_fn, steps, seed = SCENARIOS[name]
    rec, stats, n_uniforms = build_record(ns, cfg, craft, actions_fn, steps, seed)
    np.savez_compressed(os.path.join(OUT, "driving_%s.npz" % name), **rec)
    print("%-22s steps=%3d max cached pairs=%2d crashed=%2d finished=%2d dead peds=%2d uniforms=%d"
          % (name, steps, stats["maxarb"], stats["crashed"], stats["finished"], stats["dead"], n_uniforms))


if __name__ == "__main__":
    ns = load_reference()
    ns._orig_processAction = ns.DrivingEnvironment.processAction
    for n in (sys.argv[1:] or list(SCENARIOS)):
        run_scenario(ns, n)
