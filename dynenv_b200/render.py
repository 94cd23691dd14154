"""Top-down renderer of selected environments from their device state.

The reference draws with pygame + pymunk's debug draw (RoboCupEnvironment.py:584-597 drawStaticObjects,
DrivingEnvironment.py:324-354, environment_base.py:152-159 _render_internal) -- a window per environment, off the hot
path.  Here the state records of the environments to look at are copied to the host and drawn with cv2 (the library the
reference uses for its own agent-view debug image): same canvas size, same colours (field (0, 255, 0), lines white, ball
(255, 0, 0), feet by team (255, 127(1-t), 127(1+t)), posts (0, 0, 255); Driving: (50, 50, 50) ground, lane lines white /
red edges / yellow walkways, cars (0, 255, 0) or (255, 10, 10) once crashed, pedestrians (0, 255, 255) or (255, 0, 0) dead,
obstacles and buildings (200, 200, 200)), RGB order like renderMode 'memory'.  pygame is not available in this image, so
there is no pixel golden for these frames; tests check the geometry (tests/test_render.py).  The agent-view debug image
(render_agent_view) is drawn with cv2 by the reference too, and is held pixel for pixel against images its own code drew
(tests/golden/agent_view.npz).

This module is a debugging aid, not a product hot path: it needs cv2 and numpy only.
"""
import math

import numpy as np

from . import layout as L

# RoboCupEnvironment.py:26-29,73-93; Robot.py:13-20
RC_W, RC_H, RC_SIDE, RC_LINEW = 1040, 740, 70, 5
RC_PEN_RADIUS, RC_PEN_LENGTH, RC_PEN_WIDTH, RC_PEN_DIST, RC_CC_RADIUS = 5, 60, 110, 130, 75
RC_FOOT_LEN, RC_FOOT_RAD, RC_BALL_RADIUS, RC_POST_RADIUS, RC_GOAL_W = 10.0, 7.5, 5, 5, 80
# DrivingEnvironment.py:58-115; Road.py:9-33; Car.py:9-13
DRV_W, DRV_H, DRV_ROAD_W = 1750, 1000, 35
DRV_ROADS = (((875.0, 0.0), (875.0, 1000.0), 2), ((0.0, 500.0), (1750.0, 500.0), 1))
DRV_BUILDINGS = ((365.0, 200.0), (365.0, 800.0), (1385.0, 200.0), (1385.0, 800.0))
DRV_BUILD_HALF, DRV_OBST_HALF, DRV_PED_R = (400.0, 225.0), 10.0, 5


def _cv2():
    import cv2
    return cv2


def _field(rec, name, fields):
    off, cnt, kind = fields[name]
    w = np.asarray(rec[off:off + cnt])
    return w.view(np.int32) if kind == "i" else w.view(np.float32)


def _pt(x, y):
    return int(round(float(x))), int(round(float(y)))


def robocup_lines():
    """the 11 field lines (RoboCupEnvironment.py:153-196)"""
    W, H, s, pw, pl = RC_W, RC_H, RC_SIDE, RC_PEN_WIDTH, RC_PEN_LENGTH
    return [((s, s), (s, H - s)), ((W - s, s), (W - s, H - s)), ((s, s), (W - s, s)), ((s, H - s), (W - s, H - s)),
            ((W / 2, s), (W / 2, H - s)),
            ((s, H / 2 - pw), (s + pl, H / 2 - pw)), ((s, H / 2 + pw), (s + pl, H / 2 + pw)), ((s + pl, H / 2 - pw), (s + pl, H / 2 + pw)),
            ((W - s - pl, H / 2 - pw), (W - s, H / 2 - pw)), ((W - s - pl, H / 2 + pw), (W - s, H / 2 + pw)),
            ((W - s - pl, H / 2 - pw), (W - s - pl, H / 2 + pw))]


def render_robocup(record, nPlayers):
    """record: the uint32 / int32 state words of ONE RoboCup environment (BatchedRoboCup.get_state()[e]).  Returns the
    top-down frame uint8 [740, 1040, 3], RGB."""
    cv2 = _cv2()
    rec = np.ascontiguousarray(np.asarray(record)).view(np.uint32)
    F = L.RC_FIELDS
    img = np.zeros((RC_H, RC_W, 3), np.uint8)
    img[:] = (0, 255, 0)
    white = (255, 255, 255)
    for a, b in robocup_lines():
        cv2.line(img, _pt(*a), _pt(*b), white, RC_LINEW)
    for c in ((RC_W // 2, RC_H // 2), (RC_SIDE + RC_PEN_DIST, RC_H // 2), (RC_W - RC_SIDE - RC_PEN_DIST, RC_H // 2)):
        cv2.circle(img, c, RC_PEN_RADIUS * 2, white, -1)
    cv2.circle(img, (RC_W // 2, RC_H // 2), RC_CC_RADIUS * 2, white, RC_LINEW)
    px = _field(rec, "px", F).astype(np.float64) + _field(rec, "pxLo", F)
    py = _field(rec, "py", F).astype(np.float64) + _field(rec, "pyLo", F)
    ang = _field(rec, "angle", F).astype(np.float64) + _field(rec, "angleLo", F)
    P = 2 * min(int(nPlayers), 5)
    for r in range(P):  # feet 2r (left) and 2r + 1 (right): segments of half length 10 and radius 7.5 along the foot's x axis
        team = 1 if r < P // 2 else -1
        col = (255, int(127 * (1 - team)), int(127 * (1 + team)))
        for f in (2 * r, 2 * r + 1):
            c, s = math.cos(ang[f]), math.sin(ang[f])
            a = (px[f] - RC_FOOT_LEN * c, py[f] - RC_FOOT_LEN * s)
            b = (px[f] + RC_FOOT_LEN * c, py[f] + RC_FOOT_LEN * s)
            cv2.line(img, _pt(*a), _pt(*b), col, int(2 * RC_FOOT_RAD))
    cv2.circle(img, _pt(px[20], py[20]), RC_BALL_RADIUS, (255, 0, 0), -1)
    for x in (RC_SIDE, RC_W - RC_SIDE):
        for y in (RC_H / 2 - RC_GOAL_W, RC_H / 2 + RC_GOAL_W):
            cv2.circle(img, _pt(x, y), RC_POST_RADIUS, (0, 0, 255), -1)
    return img


def _box(cv2, img, x, y, hw, hh, ang, col):
    c, s = math.cos(ang), math.sin(ang)
    pts = np.array([_pt(x + dx * c - dy * s, y + dx * s + dy * c) for dx, dy in ((-hw, -hh), (hw, -hh), (hw, hh), (-hw, hh))], np.int32)
    cv2.fillConvexPoly(img, pts, col)


def render_driving(record, nPlayers):
    """record: the state words of ONE Driving environment.  Returns uint8 [1000, 1750, 3], RGB."""
    cv2 = _cv2()
    rec = np.ascontiguousarray(np.asarray(record)).view(np.uint32)
    F = L.DRV_FIELDS
    img = np.zeros((DRV_H, DRV_W, 3), np.uint8)
    img[:] = (50, 50, 50)
    for (p0, p1, nL) in DRV_ROADS:
        d = np.array(p1) - np.array(p0)
        d = d / np.linalg.norm(d)
        n = np.array([-d[1], d[0]])  # direction rotated by +pi/2 (Road.py:27-28)
        for side in (1, -1):
            off = n * (nL + 1) * DRV_ROAD_W * side
            cv2.line(img, _pt(*(np.array(p0) + off)), _pt(*(np.array(p1) + off)), (255, 255, 0), 1)
        for i in range(-nL, nL + 1):
            off = n * i * DRV_ROAD_W
            col, th = ((255, 0, 0), 1) if abs(i) == nL else ((255, 255, 255), 2 if i == 0 else 1)
            cv2.line(img, _pt(*(np.array(p0) + off)), _pt(*(np.array(p1) + off)), col, th)
    for (bx, by) in DRV_BUILDINGS:
        _box(cv2, img, bx, by, DRV_BUILD_HALF[0], DRV_BUILD_HALF[1], 0.0, (200, 200, 200))
    n_obs, n_ped = int(_field(rec, "nObs", F)[0]), int(_field(rec, "nPed", F)[0])
    ox, oy = _field(rec, "opx", F), _field(rec, "opy", F)
    for k in range(n_obs):
        _box(cv2, img, ox[k], oy[k], DRV_OBST_HALF, DRV_OBST_HALF, 0.0, (200, 200, 200))
    ppx = _field(rec, "ppx", F).astype(np.float64) + _field(rec, "ppxLo", F)
    ppy = _field(rec, "ppy", F).astype(np.float64) + _field(rec, "ppyLo", F)
    pfl = _field(rec, "pflags", F)
    for k in range(n_ped):
        dead = bool(pfl[k] & L.DRV_PF_DEAD)
        cv2.circle(img, _pt(ppx[k], ppy[k]), DRV_PED_R, (255, 0, 0) if dead else (0, 255, 255), -1)
    cpx = _field(rec, "cpx", F).astype(np.float64) + _field(rec, "cpxLo", F)
    cpy = _field(rec, "cpy", F).astype(np.float64) + _field(rec, "cpyLo", F)
    hi, lo = _field(rec, "cang", F).view(np.uint32).astype(np.uint64), _field(rec, "cangLo", F).view(np.uint32).astype(np.uint64)
    cang = ((hi << np.uint64(32)) | lo).view(np.float64)  # float64 car angles: high / low words
    ctype, cfl = _field(rec, "ctype", F), _field(rec, "cflags", F)
    for k in range(min(int(nPlayers), L.DRV_MAXCARS)):
        t = int(ctype[k])
        crashed = bool(cfl[k] & L.DRV_CF_CRASHED)
        # Car.py:9-13,49: box of half length 10 + 5 t along the heading and half width 5 + t
        _box(cv2, img, cpx[k], cpy[k], 10.0 + 5.0 * t, 5.0 + t, float(cang[k]), (255, 10, 10) if crashed else (0, 255, 0))
    return img


# RoboCupEnvironment.py:1472-1475: canvas of the agent view (H = W // 2 - 50 rows above and below the observer, who sits at
# column xOffs looking along +x); Robot.py:16 fieldOfView
AV_H, AV_W, AV_XOFFS, AV_FOV = RC_W // 2 - 50, RC_W // 2, 150, math.pi / 4
_AV_LISTS = (("RC_R_CROSS", "RC_CAP_CROSS", 3, (255, 255, 255)), ("RC_R_FCROSS", "RC_CAP_FCROSS", 4, (255, 255, 255)),
             ("RC_R_POST", "RC_CAP_POST", 2, (255, 0, 0)), ("RC_R_ROBOT", "RC_CAP_ROBOT", 1, (0, 255, 0)),
             ("RC_R_BALL", "RC_CAP_BALL", 0, (0, 0, 255)))


def render_agent_view(raw_record, dim=None):
    """The reference's agent-view debug image (getAgentVision, RoboCupEnvironment.py:1469-1516: what ONE observer detects
    at one tick, in its head frame, drawn from the observer outwards) from a raw detection record -- one
    [RC_RAW_WORDS] slice of DynEnvRoboCupOut.raw (BatchedRoboCup(..., observationType=IMAGE).out.raw[e, t, p]): the lists
    the reference draws, after noise, false negatives and false positives.  Returns uint8 [940, 1040, 3] in the reference's
    channel order.

    The reference dims Partial / Distant sightings (colour 127 instead of 255); the record does not carry the sighting
    type, so everything is drawn bright unless `dim` (bool [RC_RAW_WORDS], True at the first word of a row) says otherwise.
    A float64 record holds its counts as numbers (tests/golden/agent_view.npz); float32 / uint32 words hold them as int32."""
    cv2 = _cv2()
    rec = np.ascontiguousarray(np.asarray(raw_record))
    if rec.dtype == np.float64:
        val, cnt = rec, rec.astype(np.int64)
    else:
        val, cnt = rec.view(np.float32).astype(np.float64), rec.view(np.int32)
    H, xo = AV_H, AV_XOFFS
    img = np.zeros((2 * H, 2 * AV_W, 3), np.uint8)

    def shade(col, off):
        return tuple(c // 2 for c in col) if dim is not None and dim[off] else col

    for a in (AV_FOV, -AV_FOV):  # the two edges of the field of view
        cv2.line(img, (xo, H), (int(xo + math.cos(a) * 1000), int(H - math.sin(a) * 1000)), (255, 255, 0))
    for i in range(int(cnt[L.RC_R_COUNTS + 5])):
        o = L.RC_R_LINE + 4 * i
        cv2.line(img, (int(xo + val[o]), int(-val[o + 1] + H)), (int(xo + val[o + 2]), int(-val[o + 3] + H)), shade((255, 255, 255), o), RC_LINEW)
    if cnt[L.RC_R_CIRCLE_SEEN]:
        o = L.RC_R_CIRCLE
        cv2.circle(img, (int(xo + val[o]), int(-val[o + 1] + H)), int(val[o + 2]), shade((255, 0, 255), o), RC_LINEW)
    for off_name, cap_name, k, col in _AV_LISTS:
        off = getattr(L, off_name)
        for i in range(min(int(cnt[L.RC_R_COUNTS + k]), getattr(L, cap_name))):
            o = off + 3 * i
            cv2.circle(img, (int(xo + val[o]), int(-val[o + 1] + H)), int(val[o + 2]), shade(col, o), -1)
    return img


def render_env(sim, env_index):
    """top-down RGB frame of environment `env_index` of a BatchedRoboCup / BatchedDriving simulator"""
    rec = sim.get_state()[int(env_index)].cpu().numpy()
    if hasattr(sim, "P"):
        return render_robocup(rec, sim.nPlayers)
    return render_driving(rec, sim.nPlayers)
