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
