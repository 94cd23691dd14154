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
    return np.load(os.path.join(GOLDEN, "%s%s.npz" % (prefix, name)))


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
