"""Generates tests/golden/image_*.npz: the 8 x 480 x 640 IMAGE observations the REFERENCE's own Python
(/root/reference/DynEnv, unmodified, real cv2) returns from reset() -- the only call that works in IMAGE mode, step()
raises TypeError in processSeens (RoboCupEnvironment.py:1571).  Run here only: python tests/golden/gen_golden_image.py

Stored per scenario: configuration, where the reset's uniforms sit on the PCG64 tape (UniformTape(seed)), the state
record after the reset, and the images of every agent at the nTimeSteps = 5 initial observations, bit-packed
([5][P][8][480][20] uint32, bit x % 32 of word x // 32 = pixel x)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import raster_oracle, ref_state  # noqa: E402
from oracle.ref_loader import UniformTape, load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
# name -> (nPlayers, noiseType, noiseMagnitude, randomInit, seed)
SCENARIOS = {"5v5_realistic": (5, 1, 2.0, False, 51), "3v3_random_init": (3, 0, 3.0, True, 52), "2v2_clean": (2, 1, 0.0, True, 53)}


def build_record(ns, nPlayers, noiseType, mag, rinit, seed):
    """-> the fixture's arrays (also built on the spot for other seeds by tests/test_oracle_live.py)"""
    tape = UniformTape(seed)
    ns.RoboCupEnvironment.randomInit = rinit
    try:
        with tape.installed():
            env = ns.RoboCupEnvironment(nPlayers, observationType=ns.ObservationType.IMAGE, noiseType=ns.NoiseType(noiseType),
                                        noiseMagnitude=mag)
        tape.take_log()
        with tape.installed():
            obs = env.reset()
    finally:
        ns.RoboCupEnvironment.randomInit = False
    reset_tape = tape.take_log()
    P = 2 * nPlayers
    imgs = np.stack([np.stack([raster_oracle.pack(obs[t][a]) for a in range(P)]) for t in range(len(obs))])
    assert all(set(np.unique(obs[t][a])) <= {0.0, 1.0} for t in range(len(obs)) for a in range(P))
    return dict(images=imgs, tape_seed=np.array(seed),
                tape_reset_off=np.array([tape.count - len(reset_tape), tape.count]), tape_reset_sum=np.array(reset_tape.sum()),
                state_reset=ref_state.robocup_get_state(env),
                cfg=np.array([nPlayers, 2, noiseType, mag, 1, 0, int(rinit)], dtype=np.float64))


def run(ns, name):
    rec = build_record(ns, *SCENARIOS[name])
    imgs, P = rec["images"], rec["images"].shape[1]
    np.savez_compressed(os.path.join(OUT, "image_%s.npz" % name), **rec)
    print("%-18s agents=%d set pixels per image: %s" % (name, P, [int(raster_oracle.unpack(imgs[0, a]).sum()) for a in range(P)]))


if __name__ == "__main__":
    ns = load_reference()
    for n in (sys.argv[1:] or list(SCENARIOS)):
        run(ns, n)
