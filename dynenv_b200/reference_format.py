"""The reference's ragged, nested observation format built from the padded-with-counts records.

The batched simulator keeps observations as device tensors; these helpers materialise, for chosen
environments, exactly what DynEnv's `step()` / `reset()` return (RoboCupEnvironment.py:1537-1561 for
PARTIAL, :440-443 for FULL), for drop-in use at small batch sizes and for the parity tests.  Inputs are
numpy arrays (host copies of the views of `RoboCupOutputs`)."""
import numpy as np

from . import layout as L

_TYPES = ["ball", "robot", "post", "cross", "fcross", "line"]


def _rows(arr, n):
    """first n rows as float32 [n, feat]; an empty list is np.array([]) like the reference's"""
    if n == 0:
        return np.array([], dtype=np.float32)
    return np.array(arr[:n], dtype=np.float32)


def partial_observation(views, e, t, a, P):
    """observation of agent a at tick t of environment e: ((balls, robots), (posts, crosses, fieldCrosses,
    lines), (numLandMarks, robotsSeen, ballsSeen))"""
    c = views["counts"][e, t, a]
    lists = [_rows(views[k][e, t, a], int(c[i])) for i, k in enumerate(_TYPES)]
    seens = (int(views["num_landmarks"][e, t, a]), np.array(views["robots_seen"][e, t, a, :P - 1], dtype=np.uint8),
             bool(views["balls_seen"][e, t, a]))
    return (lists[0], lists[1]), (lists[2], lists[3], lists[4], lists[5]), seens


def full_observation(views, e, t, a, P):
    """get_full_obs entry: ([ball[1,4], others[P-1,6]], [self[1,8]], (1, 1, 1))"""
    return ([np.array(views["full_ball"][e, t, a][None], dtype=np.float32),
             np.array(views["full_others"][e, t, a, :P - 1], dtype=np.float32)],
            [np.array(views["full_self"][e, t, a][None], dtype=np.float32)], (1, 1, 1))


def unpack_image(words):
    """uint32 / int32 bit planes [..., 480, 20] -> float64 [..., 480, 640] of 0.0 / 1.0, the dtype the reference returns
    (np.zeros((4, 480, 640)) planes drawn with colour 1, RoboCupEnvironment.py:1321-1322,1533-1534)"""
    w = np.ascontiguousarray(words).view(np.uint32)
    bits = (w[..., None] >> np.arange(32, dtype=np.uint32)) & 1
    return bits.reshape(w.shape[:-1] + (w.shape[-1] * 32,)).astype(np.float64)


def image_observations(images, e, P):
    """IMAGE mode: [[img(8, 480, 640) float64 for every agent] for every tick] (RoboCupEnvironment.py:488-492,1533-1534)"""
    return [[unpack_image(images[e, t, a]) for a in range(P)] for t in range(L.RC_NTICK)]


def observations(views, e, P, full):
    T = L.RC_NTICK
    fn = full_observation if full else partial_observation
    return [[fn(views, e, t, a, P) for a in range(P)] for t in range(T)]


def info_dict(views, e, P):
    """info = {'Full State': getFullState(), 'Recon States': [getFullState(robot) ...]} (:511-512)"""
    info = {"Full State": [np.array(views["state_robots"][e, :P], dtype=np.float32), np.array(views["state_ball"][e], dtype=np.float32)],
            "Recon States": [(np.array(views["recon_ball"][e, a][None], dtype=np.float32),
                              np.array(views["recon_self"][e, a][None], dtype=np.float32),
                              np.array(views["recon_others"][e, a, :P - 1], dtype=np.float32)) for a in range(P)]}
    if bool(views["done"][e]):  # :516-521
        info["episode_r"] = np.array(views["episode_r"][e, :P], dtype=np.float64)
        info["episode_p_r"] = np.array(views["episode_p_r"][e, :P], dtype=np.float64)
        info["episode_o_r"] = np.array(views["episode_o_r"][e, :P], dtype=np.float64)
        info["episode_g"] = [int(views["episode_g"][e, 0]), int(views["episode_g"][e, 1])]
    return info


VIEW_KEYS_PARTIAL = _TYPES + ["counts", "num_landmarks", "robots_seen", "balls_seen"]
VIEW_KEYS_FULL = ["full_ball", "full_self", "full_others"]
VIEW_KEYS_INFO = ["rewards", "done", "state_robots", "state_ball", "recon_ball", "recon_self", "recon_others", "episode_r",
                  "episode_p_r", "episode_o_r", "episode_g"]


def host_views(out, full, env_ids=None):
    """numpy copies of the named views of a RoboCupOutputs (optionally only some environments)"""
    keys = (VIEW_KEYS_FULL if full else VIEW_KEYS_PARTIAL) + VIEW_KEYS_INFO
    v = {}
    for k in keys:
        t = getattr(out, k)
        if env_ids is not None:
            t = t[env_ids]
        v[k] = t.detach().cpu().numpy()
    return v


# ---------------------------------------------------------------------------------------------- Driving
# DrivingEnvironment.getAgentVision returns ((cars[.,7], obstacles[.,6], pedestrians[.,2]), (self[1,9], lanes[.,4]), (1,1,1))
# (DrivingEnvironment.py:960-977); get_full_obs ([others[N-1,7], obstacles[.,4], pedestrians[.,2]], [self[1,9], lanes[8,5]], (1,1,1))
# (:121-124).  Inputs: numpy copies of the views of `DrivingOutputs`.
DRV_VIEW_KEYS = ["self_", "car", "obs", "ped", "lane", "counts", "rewards", "done", "state_cars", "state_obs", "state_ped", "n_obs",
                 "n_ped", "goals", "episode_r", "episode_p_r", "episode_g"]


def drv_host_views(out, env_ids=None):
    v = {}
    for k in DRV_VIEW_KEYS:
        t = getattr(out, k)
        if env_ids is not None:
            t = t[env_ids]
        v[k] = t.detach().cpu().numpy()
    return v


def drv_lanes_full():
    """the constant lane rows of getFullState (DrivingEnvironment.py:690-696), including the reference's negative
    Python index: row i holds the geometry of lane offset (i - nLanes) mod (2 nLanes + 1) - nLanes"""
    rows = []
    snx, sny = 0.5 / 1800.0, 0.5 / 1100.0
    for (nL, p0, p1, nrm) in ((2, (875.0, 0.0), (875.0, 1000.0), (-1.0, 0.0)), (1, (0.0, 500.0), (1750.0, 500.0), (0.0, 1.0))):
        cnt = 2 * nL + 1
        for i in range(-nL, nL + 1):
            off = ((i - nL) % cnt) - nL
            l0 = (p0[0] + nrm[0] * off * 35.0, p0[1] + nrm[1] * off * 35.0)
            l1 = (p1[0] + nrm[0] * off * 35.0, p1[1] + nrm[1] * off * 35.0)
            rows.append([l0[0] * snx * 2, l0[1] * sny * 2, l1[0] * snx * 2, l1[1] * sny * 2,
                         1 if abs(i) == nL else (-1 if i == 0 else 0)])
    return np.array(rows, dtype=np.float32)


def drv_observations(views, e, N, full):
    """[[obs of car 0, ... car N-1]] (nTimeSteps = 1)"""
    out = []
    for a in range(N):
        c = views["counts"][e, a]
        cars = _rows(views["car"][e, a], int(c[0]))
        obst = _rows(views["obs"][e, a][:, :4] if full else views["obs"][e, a], int(c[1]))
        peds = _rows(views["ped"][e, a], int(c[2]))
        lanes = _rows(views["lane"][e, a] if full else views["lane"][e, a][:, :4], int(c[3]))
        self_ = np.array(views["self_"][e, a][None], dtype=np.float32)
        if full:
            out.append(([cars, obst, peds], [self_, lanes], (1, 1, 1)))
        else:
            out.append(((cars, obst, peds), (self_, lanes), (1, 1, 1)))
    return [out]


def drv_full_state(views, e, N, agent=None):
    """getFullState(agent) (DrivingEnvironment.py:686-747)"""
    cars = np.array(views["state_cars"][e, :N], dtype=np.float32)
    tail = [_rows(views["state_obs"][e], int(views["n_obs"][e])), _rows(views["state_ped"][e], int(views["n_ped"][e])), drv_lanes_full()]
    if agent is None:
        return [cars] + tail
    self_ = np.concatenate([cars[agent, :6], views["goals"][e, agent], cars[agent, 6:7]]).astype(np.float32)[None]
    return [self_, np.delete(cars, agent, axis=0)] + tail


def drv_info_dict(views, e, N):
    info = {"Full State": drv_full_state(views, e, N), "Recon States": [drv_full_state(views, e, N, a) for a in range(N)]}
    if bool(views["done"][e]):  # DrivingEnvironment.py:310-316
        info["episode_r"] = np.array(views["episode_r"][e, :N], dtype=np.float64)
        info["episode_p_r"] = np.array(views["episode_p_r"][e, :N], dtype=np.float64)
        info["episode_o_r"] = [0, ] * N
        info["episode_g"] = [int(views["episode_g"][e, 0]), int(views["episode_g"][e, 1])]
    return info
