"""CPU: the Driving C oracle (oracle/drv_oracle.c) against the golden vectors produced by the reference's
own Python (tests/golden/gen_golden_driving.py).  The oracle replays the uniforms the reference consumed
(tape mode), so noisy observations, pedestrian walks and crashes must agree: states to 1e-7 relative (float64
both sides), observations bit-exact after the reference's final float32 cast, counts / flags exact."""
import os

import numpy as np
import pytest

from dynenv_b200 import layout as L
from helpers import GOLDEN, golden_names, rel_err
from oracle.oracle import DrvOracle
from oracle.ref_loader import UniformTape

NAMES = golden_names("driving_")
STATE_TOL = 1e-7
OBS_KEYS = ["self", "car", "obs", "ped", "lane"]


def drv_cfg(g):
    c = g["cfg"]
    return dict(nPlayers=int(c[0]), obsType=int(c[1]), noiseType=int(c[2]), noiseMag=float(c[3]), continuous=bool(c[4]))


def drv_field_errors(wa, wb, tol):
    out = []
    for name, (off, cnt, kind) in L.DRV_FIELDS.items():
        e = rel_err(wa[..., off:off + cnt], wb[..., off:off + cnt])
        if e > tol:
            out.append((name, e))
    return out


def _check_obs(ref, out, where):
    N = ref["counts"].shape[0]
    assert np.array_equal(ref["counts"], out["counts"][:N]), (where, ref["counts"].tolist(), out["counts"][:N].tolist())
    for k in OBS_KEYS:
        got = out[k][:N].astype(np.float32)
        assert rel_err(ref[k], got) < 5e-7, (where, k, rel_err(ref[k], got))


def test_driving_goldens_exist():
    assert len(NAMES) >= 10


@pytest.mark.parametrize("name", NAMES)
def test_driving_oracle_matches_reference(name):
    check_driving_oracle_against(dict(np.load(os.path.join(GOLDEN, "driving_%s.npz" % name))))


def check_driving_oracle_against(g):
    """g: a fixture of tests/golden/gen_golden_driving.py (committed, or built on the spot: tests/test_oracle_live.py)"""
    cfg = drv_cfg(g)
    seed = int(g["tape_seed"])
    orc = DrvOracle(cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"], cfg["continuous"])
    tape_off = g["tape_off"]
    stream = UniformTape.regenerate(seed, 0, tape_off[-1])
    r0, r1 = g["tape_reset_off"]
    assert abs(stream[r0:r1].sum() - float(g["tape_reset_sum"])) < 1e-9
    orc.set_tape(stream[r0:r1])
    orc.reset()
    out = orc.reset_obs()
    assert orc.tape_pos() == r1 - r0 and not orc.tape_underflow(), (orc.tape_pos(), r1 - r0)
    assert not drv_field_errors(orc.get_state(), g["state_reset"], 1e-12)
    _check_obs({k: g["reset_" + k] for k in OBS_KEYS + ["counts"]}, out, "reset")
    for step in range(len(g["actions"])):
        lo, hi = tape_off[step], tape_off[step + 1]
        # one-step parity from the reference's own previous state
        orc.set_state(g["states"][step - 1] if step else g["state0"])
        orc.set_tape(stream[lo:hi])
        out = orc.step(g["actions"][step])
        assert orc.tape_pos() == hi - lo and not orc.tape_underflow(), (step, orc.tape_pos(), hi - lo)
        bad = drv_field_errors(orc.get_state(), g["states"][step], STATE_TOL)
        assert not bad, (step, bad)
        assert rel_err(out["rewards"], g["rewards"][step]) < STATE_TOL, step
        assert bool(out["done"]) == bool(g["dones"][step])
        assert rel_err(g["fs_cars"][step], out["state_cars"].astype(np.float32)) < 5e-7
        _check_obs({k: g["obs_" + k][step] for k in OBS_KEYS + ["counts"]}, out, step)
    assert not orc.arb_overflow()
