"""CPU, only where the reference tree exists (this container, not the GPU box): the C oracles against the reference's own
Python on scenarios that are NOT among the committed fixtures -- other seeds, team sizes and noise settings, built on the
spot by the fixture generators (tests/golden/gen_golden*.py build_record) and held to the same bar as the committed ones
(tests/test_oracle_golden*.py: every uniform consumed in the same order, states to 1e-7 relative, observations equal
after the float32 cast, counts / flags / events exact).  The committed fixtures pin the oracle on the GPU box; these show
that it was not fitted to them."""
import os
import sys

import pytest

from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs the reference tree (/root/reference)")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

# (nPlayers, obs, noise type, magnitude, canFall, allowHeadTurn, randomInit, steps, seed, hand-made start state or None)
RC_LIVE = [(5, 1, 1, 2.0, True, False, False, 24, 1001, None), (4, 1, 1, 4.0, True, True, True, 20, 1002, None),
           (2, 1, 0, 1.5, True, False, True, 30, 1003, None), (3, 0, 1, 0.0, False, False, True, 30, 1004, None),
           (5, 1, 1, 5.0, True, False, True, 16, 1005, None),
           # contact-rich starts of the fixture generator under other seeds / noise: robots walking into each other (falls,
           # pushing penalties, cached pairs), two robots face to face, a robot at a goal post
           (5, 1, 1, 3.0, True, False, False, 30, 1006, "scrum"), (5, 1, 0, 1.0, True, False, False, 30, 1007, "scrum"),
           (2, 1, 1, 4.0, True, False, False, 16, 1008, "face_to_face"), (1, 1, 1, 2.0, True, False, False, 16, 1009, "robot_at_post"),
           # a whole 240-step episode of random play: 74 fallen-robot-steps, 40 penalised-robot-steps, the episode end
           (5, 1, 1, 2.0, True, False, True, 240, 1010, None)]
# (nPlayers, obs, noise type, magnitude, continuous, steps, seed, hand-made start state or None)
DRV_LIVE = [(10, 1, 1, 2.0, False, 24, 2001, None), (7, 1, 0, 4.0, False, 24, 2002, None), (10, 1, 1, 3.0, True, 24, 2003, None),
            (3, 0, 1, 0.0, False, 30, 2004, None),
            (6, 1, 1, 4.0, False, 10, 2005, "car_car"), (5, 1, 0, 2.0, False, 12, 2006, "car_ped"), (4, 1, 1, 1.0, False, 10, 2007, "car_static"),
            # 300 steps of ten cars at full throttle: 8 crash, 9 reach their goal, a pedestrian is run over
            (10, 1, 1, 2.0, False, 300, 2011, "accelerate_10")]


@pytest.fixture(scope="module")
def ns():
    return load_reference()


def _ident(c):
    return "seed%d%s" % (c[-2], "_" + c[-1] if c[-1] else "")


@pytest.mark.parametrize("case", RC_LIVE, ids=_ident)
def test_robocup_oracle_on_a_fresh_scenario(ns, case):
    import gen_golden as G
    from oracle.oracle import pad_words
    from test_oracle_golden import check_oracle_against
    n, obs, noise, mag, fall, head, rinit, steps, seed, start = case
    cfg = G._cfg(n, obs=obs, noise=noise, mag=mag, canFall=fall, head=head, randomInit=rinit)
    craft, actions = None, (lambda s, P, g: G._rand_actions(s, P, g, True)) if head else G._rand_actions
    if start:  # the start state and the constant actions of the committed scenario of that name
        craft, actions = G.SCENARIOS[start][1], G.SCENARIOS[start][2]
    try:
        rec, stats = G.build_record(ns, cfg, craft, actions, steps, seed)
    finally:
        ns.RoboCupEnvironment.canFall, ns.RoboCupEnvironment.randomInit = True, False  # the class defaults
    if start:
        assert stats["maxarb"] > 0  # the bodies did touch
    if steps >= 240:
        assert stats["fallen"] > 20 and stats["penal"] > 10 and rec["dones"][-1]
    for k in ("state0", "states", "state_reset"):
        rec[k] = pad_words(rec[k])
    check_oracle_against(rec)


@pytest.mark.parametrize("case", DRV_LIVE, ids=_ident)
def test_driving_oracle_on_a_fresh_scenario(ns, case):
    import gen_golden_driving as G
    from test_oracle_golden_driving import check_driving_oracle_against
    n, obs, noise, mag, cont, steps, seed, start = case
    if not hasattr(ns, "_orig_processAction"):
        ns._orig_processAction = ns.DrivingEnvironment.processAction
    cfg = G._cfg(n, obs=obs, noise=noise, mag=mag, continuous=cont)
    craft, actions = None, (lambda s, N, g: G._rand_actions(s, N, g, True)) if cont else G._rand_actions
    if start:
        craft, actions = G.SCENARIOS[start][1], G.SCENARIOS[start][2]
    try:
        rec, stats, _ = G.build_record(ns, cfg, craft, actions, steps, seed)
    finally:
        ns.DrivingEnvironment.processAction = ns._orig_processAction
    if start:
        assert stats["maxarb"] > 0 or stats["crashed"] > 0 or stats["dead"] > 0
    check_driving_oracle_against(rec)


# (nPlayers, noise type, magnitude, randomInit, seed)
IMG_LIVE = [(4, 1, 3.0, True, 3001), (1, 0, 2.0, False, 3002), (5, 1, 5.0, True, 3003)]


@pytest.mark.parametrize("case", IMG_LIVE, ids=lambda c: "seed%d" % c[-1])
def test_raster_oracle_on_fresh_images(ns, case):
    """IMAGE mode: every pixel of the 8 x 480 x 640 images the reference returns from reset() (5 ticks x 2 nPlayers robots)"""
    import gen_golden_image as G
    from test_oracle_golden_image import check_raster_oracle_against
    check_raster_oracle_against(G.build_record(ns, *case))
