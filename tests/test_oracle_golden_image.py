"""CPU: the IMAGE raster oracle (oracle/raster_oracle.py over the raw detection lists of oracle/rc_oracle.c) against
the images the reference's own Python returned from reset() in IMAGE mode (tests/golden/gen_golden_image.py).  The
oracle replays the reference's uniforms (tape mode): every pixel of every channel must agree."""
import glob
import os

import numpy as np
import pytest

from helpers import GOLDEN, golden_cfg
from oracle import raster_oracle
from oracle.oracle import RcOracle
from oracle.ref_loader import UniformTape

NAMES = sorted(os.path.basename(p)[len("image_"):-4] for p in glob.glob(os.path.join(GOLDEN, "image_*.npz")))


def test_image_goldens_exist():
    assert len(NAMES) >= 3


@pytest.mark.parametrize("name", NAMES)
def test_raster_oracle_matches_reference(name):
    check_raster_oracle_against(dict(np.load(os.path.join(GOLDEN, "image_%s.npz" % name))))


def check_raster_oracle_against(g):
    """g: a fixture of tests/golden/gen_golden_image.py (committed, or built on the spot: tests/test_oracle_live.py)"""
    cfg = golden_cfg(g)
    P = 2 * cfg["nPlayers"]
    orc = RcOracle(cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"], cfg["allowHeadTurn"], cfg["flags"])
    r0, r1 = g["tape_reset_off"]
    stream = UniformTape.regenerate(int(g["tape_seed"]), 0, r1)
    assert abs(stream[r0:r1].sum() - float(g["tape_reset_sum"])) < 1e-9
    orc.set_tape(stream[r0:r1])
    orc.capture_raw()
    orc.reset()
    orc.reset_obs()
    assert orc.tape_pos() == r1 - r0 and not orc.tape_underflow(), (orc.tape_pos(), r1 - r0)
    ref = g["images"]
    assert ref.shape[:2] == (5, P)
    for t in range(5):
        for a in range(P):
            got = raster_oracle.pack(raster_oracle.render(orc.raw_vision(t, a)))
            diff = int(np.unpackbits((got ^ ref[t, a]).view(np.uint8)).sum())
            assert diff == 0, (t, a, diff)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    img = (rng.random((8, 480, 640)) < 0.1).astype(np.float64)
    w = raster_oracle.pack(img)
    assert w.shape == (8, 480, 20) and w.dtype == np.uint32
    assert np.array_equal(raster_oracle.unpack(w), img.astype(np.uint8))
    assert (w[0, 0, 0] >> 5) & 1 == int(img[0, 0, 5])
