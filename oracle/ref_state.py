"""State extraction / injection for the REFERENCE env objects running on the shim.

TEST INFRASTRUCTURE ONLY.  Maps a reference RoboCupEnvironment (its Robot/Ball objects and
the shim Space's arbiter cache) to the float64 word list whose offsets are the C-ABI state
record layout (include/dynenv_b200.h), and back.  Used by tests/golden/gen_golden.py.
"""
import numpy as np

from dynenv_b200 import layout as L

F = L.RC_FIELDS


def _set(w, name, vals):
    off, cnt, _ = F[name]
    v = np.atleast_1d(np.asarray(vals, dtype=np.float64))
    w[off:off + len(v)] = v


def _get(w, name):
    off, cnt, kind = F[name]
    return w[off:off + cnt]


def _shape_id(env, shape):
    """shim shape -> C-ABI shape id (feet 0..19, ball 20, posts 21..24)."""
    for r, rob in enumerate(env.agents):
        if shape is rob.leftFoot:
            return 2 * r
        if shape is rob.rightFoot:
            return 2 * r + 1
    if shape is env.ball.shape:
        return 20
    for k, g in enumerate(env.goalposts):
        if shape is g.shape:
            return 21 + k
    raise KeyError(shape)


def _contact_hash(h):
    if h == 0:
        return 0
    (_, v1), (_, v2) = h
    return 1 + (v1 | (v2 << 1))


def robocup_get_state(env, episode=1):
    w = np.zeros(L.RC_STATE_WORDS, dtype=np.float64)
    bodies = []
    for rob in env.agents:
        bodies += [rob.leftFoot, rob.rightFoot]
    ids = list(range(len(bodies))) + [20]
    shapes = bodies + [env.ball.shape]
    for i, sh in zip(ids, shapes):
        b = sh.body
        for name, val in (("px", b._px), ("py", b._py), ("angle", b._a), ("vx", b._vx), ("vy", b._vy),
                          ("w", b._w), ("fx", b._fx), ("fy", b._fy), ("torque", b._t),
                          ("vbx", b._vbx), ("vby", b._vby), ("wb", b._wb)):
            w[F[name][0] + i] = val
        if i == 20:
            ca = cb = sh.tc
        else:
            ca, cb = sh.ta, sh.tb
        w[F["cax"][0] + i], w[F["cay"][0] + i] = ca
        w[F["cbx"][0] + i], w[F["cby"][0] + i] = cb
    absent_pivots = set(id(r.joint) for r in env.agents) - set(id(c) for c in env.space.constraints)
    for r, rob in enumerate(env.agents):
        w[F["headAngle"][0] + r] = rob.headAngle
        w[F["headMoving"][0] + r] = rob.headMoving
        w[F["moveTime"][0] + r] = rob.moveTime
        if rob.initPos is not None:
            w[F["initX"][0] + r], w[F["initY"][0] + r] = rob.initPos
        w[F["fallTime"][0] + r] = rob.fallTime
        w[F["penalTime"][0] + r] = rob.penalTime
        w[F["prevX"][0] + r], w[F["prevY"][0] + r] = rob.prevPos
        w[F["pivJx"][0] + r], w[F["pivJy"][0] + r] = rob.joint.jAcc
        w[F["rotJ"][0] + r] = rob.rotJoint.jAcc
        fl = ((L.RC_RF_KICKING if rob.kicking else 0) | (L.RC_RF_FOOT if rob.foot else 0)
              | (L.RC_RF_FALLEN if rob.fallen else 0) | (L.RC_RF_PENALIZED if rob.penalized else 0)
              | (L.RC_RF_TOUCHING if rob.touching else 0) | (L.RC_RF_MIGHTPUSH if rob.mightPush else 0)
              | (L.RC_RF_JOINT_REMOVED if rob.jointRemoved else 0)
              | (L.RC_RF_JOINT_ABSENT if id(rob.joint) in absent_pivots else 0))
        w[F["rflags"][0] + r] = fl
        w[F["fallCntr"][0] + r] = rob.fallCntr
        w[F["touchCntr"][0] + r] = rob.touchCntr
    P = len(env.agents)
    for name, src in (("episodeRewards", env.episodeRewards), ("episodePosRewards", env.episodePosRewards),
                      ("episodeObsRewards", env.episodeObsRewards)):
        w[F[name][0]:F[name][0] + P] = np.broadcast_to(np.asarray(src, dtype=np.float64), (P,))
    _set(w, "ballPrevX", env.ball.prevPos[0])
    _set(w, "ballPrevY", env.ball.prevPos[1])
    _set(w, "lastKicked", list(env.ball.lastKicked))
    _set(w, "nLastKicked", len(env.ball.lastKicked))
    _set(w, "elapsed", env.elapsed)
    _set(w, "episode", episode)
    _set(w, "ballOwned", env.ballOwned)
    _set(w, "ballFreeCntr", env.ballFreeCntr)
    _set(w, "gracePeriod", env.gracePeriod)
    _set(w, "goals", env.goals)
    _set(w, "closestID", env.closestID)
    d = np.zeros(4)
    for t in range(2):
        for k, v in enumerate(env.defenders[t]):
            d[2 * t + k] = v
    _set(w, "defenders", d)
    _set(w, "nDefenders", [len(env.defenders[0]), len(env.defenders[1])])
    _set(w, "penalTimes", env.penalTimes)
    _set(w, "stepped", 1 if env.space.curr_dt != 0.0 else 0)
    arbs = list(env.space.cached_arbiters.values())
    assert len(arbs) <= L.RC_MAXARB
    _set(w, "nArb", len(arbs))
    for k, arb in enumerate(arbs):
        a = _shape_id(env, arb.a)
        b = _shape_id(env, arb.b)
        # shim states: 0 FIRST 1 NORMAL 2 IGNORE 3 CACHED (same numbering as the ABI)
        age = env.space.stamp - arb.stamp
        w[F["arbKey"][0] + k] = a | (b << 8) | (arb.state << 16) | (age << 24)
        hs = [_contact_hash(c.hash) for c in arb.contacts] + [0, 0]
        w[F["arbCon"][0] + k] = len(arb.contacts) | (hs[0] << 8) | (hs[1] << 16)
        for ci, c in enumerate(arb.contacts[:2]):
            w[F["arbJn%d" % ci][0] + k] = c.jnAcc
            w[F["arbJt%d" % ci][0] + k] = c.jtAcc
    return w


def robocup_set_state(env, w, ns):
    """Injects a word list into a freshly reset reference env (same nPlayers)."""
    w = [float(v) for v in w]  # plain Python floats, as the reference itself would hold
    pm = ns.pymunk
    shapes = []
    for rob in env.agents:
        shapes += [rob.leftFoot, rob.rightFoot]
    ids = list(range(len(shapes))) + [20]
    shapes = shapes + [env.ball.shape]
    for i, sh in zip(ids, shapes):
        b = sh.body
        b._px, b._py, b._a = (w[F[n][0] + i] for n in ("px", "py", "angle"))
        b._vx, b._vy, b._w = (w[F[n][0] + i] for n in ("vx", "vy", "w"))
        b._fx, b._fy, b._t = (w[F[n][0] + i] for n in ("fx", "fy", "torque"))
        b._vbx, b._vby, b._wb = (w[F[n][0] + i] for n in ("vbx", "vby", "wb"))
        ca = (w[F["cax"][0] + i], w[F["cay"][0] + i])
        cb = (w[F["cbx"][0] + i], w[F["cby"][0] + i])
        if i == 20:
            sh.tc = ca
        else:
            sh.ta, sh.tb = ca, cb
    for r, rob in enumerate(env.agents):
        rob.headAngle = w[F["headAngle"][0] + r]
        rob.headMoving = w[F["headMoving"][0] + r]
        rob.moveTime = w[F["moveTime"][0] + r]
        rob.initPos = pm.Vec2d(w[F["initX"][0] + r], w[F["initY"][0] + r])
        rob.fallTime = w[F["fallTime"][0] + r]
        rob.penalTime = w[F["penalTime"][0] + r]
        rob.prevPos = pm.Vec2d(w[F["prevX"][0] + r], w[F["prevY"][0] + r])
        rob.joint.jAcc = (w[F["pivJx"][0] + r], w[F["pivJy"][0] + r])
        rob.rotJoint.jAcc = w[F["rotJ"][0] + r]
        fl = int(w[F["rflags"][0] + r])
        rob.kicking = bool(fl & L.RC_RF_KICKING)
        rob.foot = 1 if fl & L.RC_RF_FOOT else 0
        rob.fallen = bool(fl & L.RC_RF_FALLEN)
        rob.penalized = bool(fl & L.RC_RF_PENALIZED)
        rob.touching = bool(fl & L.RC_RF_TOUCHING)
        rob.mightPush = bool(fl & L.RC_RF_MIGHTPUSH)
        rob.jointRemoved = bool(fl & L.RC_RF_JOINT_REMOVED)
        present = rob.joint in env.space.constraints
        if fl & L.RC_RF_JOINT_ABSENT and present:
            env.space.remove(rob.joint)
        elif not (fl & L.RC_RF_JOINT_ABSENT) and not present:
            env.space.add(rob.joint)
        rob.fallCntr = int(w[F["fallCntr"][0] + r])
        rob.touchCntr = int(w[F["touchCntr"][0] + r])
    P = len(env.agents)
    env.episodeRewards = np.array(w[F["episodeRewards"][0]:F["episodeRewards"][0] + P])
    env.episodePosRewards = np.array(w[F["episodePosRewards"][0]:F["episodePosRewards"][0] + P])
    env.episodeObsRewards = np.array(w[F["episodeObsRewards"][0]:F["episodeObsRewards"][0] + P])
    env.ball.prevPos = pm.Vec2d(_get(w, "ballPrevX")[0], _get(w, "ballPrevY")[0])
    n = int(_get(w, "nLastKicked")[0])
    env.ball.lastKicked = [int(v) for v in _get(w, "lastKicked")[:n]]
    env.elapsed = int(_get(w, "elapsed")[0])
    env.ballOwned = int(_get(w, "ballOwned")[0])
    env.ballFreeCntr = float(_get(w, "ballFreeCntr")[0])
    env.gracePeriod = float(_get(w, "gracePeriod")[0])
    env.goals = np.array([int(v) for v in _get(w, "goals")])
    env.closestID = [int(v) for v in _get(w, "closestID")]
    nd = [int(v) for v in _get(w, "nDefenders")]
    d = [int(v) for v in _get(w, "defenders")]
    env.defenders = [d[0:nd[0]], d[2:2 + nd[1]]]
    env.penalTimes = [int(v) for v in _get(w, "penalTimes")]
    sp = env.space
    sp.curr_dt = 0.01 if int(_get(w, "stepped")[0]) else 0.0
    sp.cached_arbiters = {}
    sp.arbiters = []
    by_id = {}
    for i, sh in zip(ids, shapes):
        by_id[i] = sh
    for k, g in enumerate(env.goalposts):
        by_id[21 + k] = g.shape
    for k in range(int(_get(w, "nArb")[0])):
        key = int(w[F["arbKey"][0] + k])
        cn = int(w[F["arbCon"][0] + k])
        a, b, st, age = key & 255, (key >> 8) & 255, (key >> 16) & 255, (key >> 24) & 255
        arb = pm.Arbiter(by_id[a], by_id[b])
        arb.state = st
        arb.stamp = sp.stamp - age
        h = sp._lookup_handler(arb.a.collision_type, arb.b.collision_type)
        arb.handler = h
        arb.swapped = (h is not None and arb.a.collision_type != h.type_a)
        hs = [(cn >> 8) & 255, (cn >> 16) & 255]
        for ci in range(cn & 255):
            hv = hs[ci]
            hh = 0 if hv == 0 else ((arb.a.index, (hv - 1) & 1), (arb.b.index, ((hv - 1) >> 1) & 1))
            c = pm.Contact((0.0, 0.0), (0.0, 0.0), hh)
            c.jnAcc = w[F["arbJn%d" % ci][0] + k]
            c.jtAcc = w[F["arbJt%d" % ci][0] + k]
            arb.contacts.append(c)
        sp.cached_arbiters[(arb.a.index, arb.b.index)] = arb
        if st == 0:  # FIRST: was active in the previous substep
            sp.arbiters.append(arb)


# ---------------------------------------------------------------- Driving
DF = L.DRV_FIELDS


def _drv_shape_id(env, shape):
    """shim shape -> C-ABI shape id (buildings 0..3, cars 4..13, pedestrians 14..33, obstacles 34..53)"""
    for k, b in enumerate(env.buildings):
        if shape is b.shape:
            return k
    for k, c in enumerate(env.agents):
        if shape is c.shape:
            return L.DRV_SHAPE_CAR0 + k
    for k, p in enumerate(env.pedestrians):
        if shape is p.shape:
            return L.DRV_SHAPE_PED0 + k
    for k, o in enumerate(env.obstacles):
        if shape is o.shape:
            return L.DRV_SHAPE_OBS0 + k
    raise KeyError(shape)


def _drv_hash(h):
    if h == 0:
        return 0
    (_, v1), (_, v2) = h
    return 1 + v1 + 4 * v2


def driving_get_state(env, episode=1):
    w = np.zeros(L.DRV_STATE_WORDS, dtype=np.float64)
    for i, c in enumerate(env.agents):
        b = c.shape.body
        for name, val in (("cpx", b._px), ("cpy", b._py), ("cang", b._a), ("cvx", b._vx), ("cvy", b._vy), ("cw", b._w),
                          ("cvbx", b._vbx), ("cvby", b._vby), ("cwb", b._wb), ("cdirx", c.direction.x), ("cdiry", c.direction.y),
                          ("cprevx", c.prevPos[0]), ("cprevy", c.prevPos[1]), ("cgoalx", c.goal[0]), ("cgoaly", c.goal[1]),
                          ("ctype", c.type), ("cteam", c.team),
                          ("cflags", (L.DRV_CF_FINISHED if c.finished else 0) | (L.DRV_CF_CRASHED if c.crashed else 0)),
                          ("clane", int(c.position))):
            w[DF[name][0] + i] = val
    for i, p in enumerate(env.pedestrians):
        b = p.shape.body
        road = 0 if p.direction == env.roads[0].direction else 1
        for name, val in (("ppx", b._px), ("ppy", b._py), ("pvx", b._vx), ("pvy", b._vy), ("pvbx", b._vbx), ("pvby", b._vby),
                          ("pmoving", p.moving), ("pspeed", p.speed),
                          ("pflags", (L.DRV_PF_DEAD if p.dead else 0) | (L.DRV_PF_CROSSING if p.crossing else 0)
                           | (L.DRV_PF_BEGINCROSSING if p.beginCrossing else 0) | (L.DRV_PF_SIDE if p.side else 0)
                           | (L.DRV_PF_ROAD if road else 0))):
            w[DF[name][0] + i] = val
    for i, o in enumerate(env.obstacles):
        w[DF["opx"][0] + i], w[DF["opy"][0] + i] = o.shape.body._px, o.shape.body._py
    N = len(env.agents)
    w[DF["episodeRewards"][0]:DF["episodeRewards"][0] + N] = env.episodeRewards
    w[DF["episodePosRewards"][0]:DF["episodePosRewards"][0] + N] = env.episodePosRewards
    for name, val in (("elapsed", env.elapsed), ("episode", episode), ("allFinished", int(env.allFinished)),
                      ("nPed", len(env.pedestrians)), ("nObs", len(env.obstacles)),
                      ("stepped", 1 if env.space.curr_dt != 0.0 else 0)):
        w[DF[name][0]] = val
    # only pairs that involve a car are state: the others are ignored by the handlers (no contacts, no callbacks)
    arbs = [a for a in env.space.cached_arbiters.values()
            if any(L.DRV_SHAPE_CAR0 <= _drv_shape_id(env, s) < L.DRV_SHAPE_PED0 for s in (a.a, a.b))]
    assert len(arbs) <= L.DRV_MAXARB
    w[DF["nArb"][0]] = len(arbs)
    for k, arb in enumerate(arbs):
        a, b = _drv_shape_id(env, arb.a), _drv_shape_id(env, arb.b)
        age = env.space.stamp - arb.stamp
        w[DF["arbKey"][0] + k] = a | (b << 8) | (arb.state << 16) | (age << 24)
        hs = [_drv_hash(c.hash) for c in arb.contacts] + [0, 0]
        w[DF["arbCon"][0] + k] = len(arb.contacts) | (hs[0] << 8) | (hs[1] << 16)
        for ci, c in enumerate(arb.contacts[:2]):
            w[DF["arbJn%d" % ci][0] + k] = c.jnAcc
            w[DF["arbJt%d" % ci][0] + k] = c.jtAcc
    return w


def driving_set_state(env, w, ns):
    """Injects positions / velocities / flags of cars and pedestrians into a freshly reset reference env (same
    scene: types, goals, counts are taken from the env itself); clears the arbiter cache."""
    w = [float(v) for v in w]
    Vec2d = ns.pymunk.Vec2d
    friction_car = ns.cutils.friction_car
    friction_car_crashed = ns.cutils.friction_car_crashed
    for i, c in enumerate(env.agents):
        b = c.shape.body
        b._px, b._py, b._a = w[DF["cpx"][0] + i], w[DF["cpy"][0] + i], w[DF["cang"][0] + i]
        b._vx, b._vy, b._w = w[DF["cvx"][0] + i], w[DF["cvy"][0] + i], w[DF["cw"][0] + i]
        b._vbx = b._vby = b._wb = 0.0
        c.direction = Vec2d(w[DF["cdirx"][0] + i], w[DF["cdiry"][0] + i])
        c.prevPos = Vec2d(w[DF["cprevx"][0] + i], w[DF["cprevy"][0] + i])
        fl = int(w[DF["cflags"][0] + i])
        c.finished, c.crashed = bool(fl & L.DRV_CF_FINISHED), bool(fl & L.DRV_CF_CRASHED)
        b.velocity_func = friction_car_crashed if c.finished else friction_car
        c.shape.cache_bb()
    for i, p in enumerate(env.pedestrians):
        b = p.shape.body
        b._px, b._py = w[DF["ppx"][0] + i], w[DF["ppy"][0] + i]
        b._vx, b._vy = w[DF["pvx"][0] + i], w[DF["pvy"][0] + i]
        p.moving = w[DF["pmoving"][0] + i]
        fl = int(w[DF["pflags"][0] + i])
        p.crossing, p.beginCrossing, p.side = bool(fl & L.DRV_PF_CROSSING), bool(fl & L.DRV_PF_BEGINCROSSING), 1 if fl & L.DRV_PF_SIDE else 0
        p.shape.cache_bb()
    env.space.cached_arbiters.clear()
    env.space.arbiters = []
