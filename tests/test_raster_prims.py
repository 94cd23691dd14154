"""CPU: the rasterisation primitives of the IMAGE observation (dynenv_b200/csrc/raster_prims.cuh, the code the CUDA
raster kernel runs) compiled for the host with g++ and compared with the installed cv2 -- the library the reference
draws with (RoboCupEnvironment.py:1318-1467) -- on random lines, poly-lines, filled rectangles and filled circles,
including primitives far outside the image.  Bit-exact planes are required."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "tests", "native", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libraster_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "native", "raster_host.cpp")])
    return C.CDLL(so)


def _unpack(w):
    return ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(480, 640).astype(np.uint8)


def _coord(rng, far):
    if far:
        return int(rng.integers(-3000, 4000)), int(rng.integers(-3000, 4000))
    return int(rng.integers(-60, 700)), int(rng.integers(-60, 540))


def test_primitives_match_cv2(lib):
    rng = np.random.default_rng(7)
    for it in range(1500):
        far = rng.random() < 0.3
        x0, y0 = _coord(rng, far)
        x1, y1 = _coord(rng, far)
        if rng.random() < 0.15:
            x1 = x0
        if rng.random() < 0.15:
            y1 = y0
        t = int(rng.integers(1, 40))
        ref = np.zeros((480, 640))
        cv2.line(ref, (x0, y0), (x1, y1), 1, t)
        p = np.zeros((480, 20), np.uint32)
        lib.rp_line(p.ctypes.data, x0, y0, x1, y1, t)
        assert np.array_equal(_unpack(p), ref.astype(np.uint8)), ("line", x0, y0, x1, y1, t)
        ref = np.zeros((480, 640))
        cv2.rectangle(ref, (x0, y0), (x1, y1), 1, -1)
        p = np.zeros((480, 20), np.uint32)
        lib.rp_rect(p.ctypes.data, x0, y0, x1, y1)
        assert np.array_equal(_unpack(p), ref.astype(np.uint8)), ("rectangle", x0, y0, x1, y1)
        r = int(rng.integers(0, 300)) if rng.random() < 0.8 else int(rng.integers(0, 5))
        ref = np.zeros((480, 640))
        cv2.circle(ref, (x0, y0), r, 1, -1)
        p = np.zeros((480, 20), np.uint32)
        lib.rp_circle(p.ctypes.data, x0, y0, r)
        assert np.array_equal(_unpack(p), ref.astype(np.uint8)), ("circle", x0, y0, r)


def test_cooperative_lanes_draw_the_same_pixels(lib):
    """the CUDA kernel spreads rows / pixels of a primitive over 32 lanes (or 256 threads) in closed form: emulate the
    lanes one after the other on the host"""
    rng = np.random.default_rng(9)
    for lanes in (32, 256):
        lib.rp_set_lanes(lanes)
        try:
            for it in range(250):
                far = rng.random() < 0.3
                x0, y0 = _coord(rng, far)
                x1, y1 = _coord(rng, far)
                t = int(rng.integers(1, 60))
                ref = np.zeros((480, 640))
                cv2.line(ref, (x0, y0), (x1, y1), 1, t)
                p = np.zeros((480, 20), np.uint32)
                lib.rp_line(p.ctypes.data, x0, y0, x1, y1, t)
                assert np.array_equal(_unpack(p), ref.astype(np.uint8)), ("line", lanes, x0, y0, x1, y1, t)
                r = int(rng.integers(0, 200))
                ref = np.zeros((480, 640))
                cv2.circle(ref, (x0, y0), r, 1, -1)
                cv2.rectangle(ref, (x0, y0), (x1, y1), 1, -1)
                p = np.zeros((480, 20), np.uint32)
                lib.rp_circle(p.ctypes.data, x0, y0, r)
                lib.rp_rect(p.ctypes.data, x0, y0, x1, y1)
                assert np.array_equal(_unpack(p), ref.astype(np.uint8)), ("circle+rect", lanes, x0, y0, x1, y1, r)
        finally:
            lib.rp_set_lanes(1)


def test_polylines_match_cv2(lib):
    rng = np.random.default_rng(8)
    for it in range(150):
        n = int(rng.integers(1, 60))
        t = int(rng.integers(1, 31))
        pts = (np.cumsum(rng.integers(-15, 16, (n, 2)), 0) + np.array(_coord(rng, False))).astype(np.int32)
        ref = np.zeros((480, 640))
        cv2.polylines(ref, [pts], False, 1, t)
        p = np.zeros((480, 20), np.uint32)
        pc = np.ascontiguousarray(pts)
        lib.rp_polyline(p.ctypes.data, pc.ctypes.data, n, t)
        assert np.array_equal(_unpack(p), ref.astype(np.uint8)), ("polyline", pts.tolist(), t)


def test_compose_matches_raster_oracle(lib):
    """the whole image of a raw detection record (projection, conic fit, primitives) by the host build of
    raster_compose.cuh against the cv2-based raster oracle, for every agent and tick of the golden IMAGE scenes"""
    import glob
    from helpers import GOLDEN, golden_cfg
    from oracle import raster_oracle
    from oracle.oracle import RcOracle, raw_to_record, record_to_raw
    from oracle.ref_loader import UniformTape
    n = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, "image_*.npz"))):
        g = dict(np.load(path))
        cfg = golden_cfg(g)
        orc = RcOracle(cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"], cfg["allowHeadTurn"], cfg["flags"])
        r0, r1 = g["tape_reset_off"]
        orc.set_tape(UniformTape.regenerate(int(g["tape_seed"]), 0, r1)[r0:r1])
        orc.capture_raw()
        orc.reset()
        orc.reset_obs()
        for t in (0, 3):
            for a in range(2 * cfg["nPlayers"]):
                rec = raw_to_record(orc.raw_vision(t, a))
                ref = raster_oracle.pack(raster_oracle.render(record_to_raw(rec)))
                planes = np.zeros((8, 480, 20), np.uint32)
                lib.rp_render(rec.ctypes.data, planes.ctypes.data)
                assert np.array_equal(planes, ref), (os.path.basename(path), t, a)
                n += 1
    assert n >= 40


def test_both_integer_widths_match_cv2(lib):
    """the fixed-point routines run in 32-bit integers when every coordinate fits (thickness <= 2048, rectangle corners
    within +-8191 px) and in OpenCV's own 64 bits otherwise: primitives on both sides of the switch, and exactly at it"""
    rng = np.random.default_rng(17)
    for it in range(160):
        x0, y0 = _coord(rng, True)
        x1, y1 = _coord(rng, True)
        t = int(rng.choice([2047, 2048, 2049, 2050, 3000, 700, 1500, 5000, 32767])) if it % 2 else int(rng.integers(100, 2600))
        ref = np.zeros((480, 640))
        cv2.line(ref, (x0, y0), (x1, y1), 1, t)
        p = np.zeros((480, 20), np.uint32)
        lib.rp_line(p.ctypes.data, x0, y0, x1, y1, t)
        assert np.array_equal(_unpack(p), ref.astype(np.uint8)), ("line", x0, y0, x1, y1, t)
    edge = [8190, 8191, 8192, 8193, -8190, -8191, -8192, -8193, 20000, -20000, 1000000, -1000000]
    for it in range(300):
        x0, y0 = (int(rng.choice(edge)) if rng.random() < 0.5 else int(rng.integers(-700, 1300))), int(rng.integers(-600, 1100))
        x1, y1 = int(rng.integers(-700, 1300)), (int(rng.choice(edge)) if rng.random() < 0.5 else int(rng.integers(-600, 1100)))
        if rng.random() < 0.5:
            x0, y0, x1, y1 = x1, y1, x0, y0
        ref = np.zeros((480, 640))
        cv2.rectangle(ref, (x0, y0), (x1, y1), 1, -1)
        p = np.zeros((480, 20), np.uint32)
        lib.rp_rect(p.ctypes.data, x0, y0, x1, y1)
        assert np.array_equal(_unpack(p), ref.astype(np.uint8)), ("rectangle", x0, y0, x1, y1)


def test_segment_pattern_does_not_depend_on_its_position(lib):
    """an integer translation of a thick one-row poly-line segment translates its pixels (nothing clipped): OpenCV's integer /
    fixed-point routines have no position-dependent rounding.  (A stamp-based drawing of the centre circle rests on this; it was
    built, bit-exact, and measured slower than the exact routines per segment -- DESIGN.md section 8.)"""
    rng = np.random.default_rng(31)
    for t in range(2, 31):
        for dx in range(-7, 8):
            ref = np.zeros((480, 20), np.uint32)
            lib.rp_segment(ref.ctypes.data, 320, 100, 320 + dx, 101, t, 2)
            ref = _unpack(ref)
            ys, xs = np.nonzero(ref)
            assert xs.min() >= 288 and xs.max() <= 351 and ys.min() >= 81 and ys.max() <= 120, (t, dx)  # the stamp window
            for _ in range(3):
                xa, ya = int(rng.integers(40, 601)), int(rng.integers(20, 459))
                got = np.zeros((480, 20), np.uint32)
                lib.rp_segment(got.ctypes.data, xa, ya, xa + dx, ya + 1, t, 2)
                assert np.array_equal(_unpack(got), np.roll(np.roll(ref, ya - 100, 0), xa - 320, 1)), (t, dx, xa, ya)
