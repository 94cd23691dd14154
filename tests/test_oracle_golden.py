"""CPU: the C oracle (oracle/rc_oracle.c) against the golden vectors produced by the
reference's own Python (tests/golden/gen_golden.py).  The oracle replays the uniforms the
reference consumed (tape mode), so noisy scenarios, falls and penalties must agree too:
states to 1e-9 relative (float64 both sides), observations bit-exact after the reference's
final float32 cast, counts / flags / events exact."""
import numpy as np
import pytest

from helpers import FULL_KEYS, OBS_KEYS, SEEN_KEYS, field_errors, golden_cfg, golden_names, load_golden, rel_err
from oracle.oracle import RcOracle
from oracle.ref_loader import UniformTape

NAMES = golden_names()
# contacts amplify last-bit differences between the Python and C evaluation order
STATE_TOL = 1e-7


def _check_obs(g, prefix, out, cfg):
    keys = FULL_KEYS if cfg["obsType"] == 0 else OBS_KEYS + SEEN_KEYS
    for k in keys:
        ref = g[prefix + k]
        got = out[k]
        if ref.dtype == np.float32:
            # bit-exact after the float32 cast unless a contact amplified last-bit differences
            assert rel_err(ref, got.astype(np.float32)) < 5e-7, (prefix, k, rel_err(ref, got))
        else:
            assert np.array_equal(ref, got), (prefix, k)


def test_goldens_exist():
    assert len(NAMES) >= 15


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name):
    check_oracle_against(load_golden(name))


def check_oracle_against(g):
    """g: a fixture of tests/golden/gen_golden.py (committed, or built on the spot: tests/test_oracle_live.py)"""
    cfg = golden_cfg(g)
    seed = int(g["tape_seed"])
    P = 2 * cfg["nPlayers"]
    orc = RcOracle(cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"], cfg["allowHeadTurn"],
                   cfg["flags"])
    tape_off = g["tape_off"]
    stream = UniformTape.regenerate(seed, 0, tape_off[-1])
    # reset + initial observations
    r0, r1 = g["tape_reset_off"]
    assert abs(stream[r0:r1].sum() - float(g["tape_reset_sum"])) < 1e-9
    orc.set_tape(stream[r0:r1])
    orc.reset()
    out = orc.reset_obs()
    assert orc.tape_pos() == r1 - r0 and not orc.tape_underflow()
    _check_obs(g, "reset_", out, cfg)
    assert not field_errors(orc.get_state(), g["state_reset"], 1e-12)
    # scenario
    orc.set_state(g["state0"])
    obs_steps = set(int(s) for s in g["obs_steps"])
    for step in range(len(g["actions"])):
        lo, hi = tape_off[step], tape_off[step + 1]
        # one-step parity: start every step from the reference's own previous state, so that
        # contact dynamics cannot amplify last-bit differences over several steps
        orc.set_state(g["states"][step - 1] if step else g["state0"])
        orc.set_tape(stream[lo:hi])
        out = orc.step(g["actions"][step])
        assert orc.tape_pos() == hi - lo and not orc.tape_underflow(), (step, orc.tape_pos(), hi - lo)
        bad = field_errors(orc.get_state(), g["states"][step], STATE_TOL)
        assert not bad, (step, bad)
        assert rel_err(out["rewards"], g["rewards"][step]) < STATE_TOL, step
        assert bool(out["done"]) == bool(g["dones"][step])
        assert rel_err(g["fs_robots"][step][:P], out["state_robots"][:P].astype(np.float32)) < 5e-7
        assert rel_err(g["fs_ball"][step], out["state_ball"].astype(np.float32)) < 5e-7
        if step in obs_steps:
            _check_obs(g, "obs%d_" % step, out, cfg)
    assert not orc.arb_overflow()
