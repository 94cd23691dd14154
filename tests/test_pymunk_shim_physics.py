"""CPU: physical invariants of the restated Chipmunk2D subset (oracle/shim/pymunk) that stands in for pymunk.

pymunk itself is absent from this image, so the stand-in cannot be pinned against it ("parity
unpinned").  What CAN be checked is that it is a correct rigid-body solver of Chipmunk's published kind at the call sites
the reference uses (environment_base.py:126-128 Space, RoboCupEnvironment.py:482 / DrivingEnvironment.py:278 step,
Robot.py:58-60 joints): conservation laws in free collisions, the restitution law, the friction cone, penetration
settling onto collision_slop at the rate collision_bias prescribes, and the joint error decaying by error_bias^dt per
step -- the documented behaviour of cpSpaceStep / cpArbiter / cpPivotJoint.  These bound how wrong the stand-in (and with
it the C oracle and the CUDA kernels, which are pinned to it bit for bit / to 1e-4) can be about real Chipmunk."""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shim"))
import pymunk  # noqa: E402  (the stand-in)
from pymunk.constraint import PivotJoint, RotaryLimitJoint  # noqa: E402

DT = 0.01


def _circle(space, m, r, pos, vel=(0, 0), e=1.0, u=0.0, w=0.0):
    b = pymunk.Body(m, pymunk.moment_for_circle(m, 0, r))
    b.position, b.velocity, b.angular_velocity = pos, vel, w
    s = pymunk.Circle(b, r)
    s.elasticity, s.friction = e, u
    space.add(b, s)
    return b


def _momentum(bodies):
    p = np.zeros(2)
    L = 0.0
    for b in bodies:
        v, x = np.array(b.velocity), np.array(b.position)
        p += b.mass * v
        L += b.mass * (x[0] * v[1] - x[1] * v[0]) + b.moment * b.angular_velocity
    return p, L


@pytest.mark.parametrize("e1,e2,m2", [(1.0, 1.0, 1.0), (0.9, 0.5, 3.0), (0.3, 0.3, 0.25), (0.0, 1.0, 10.0)])
def test_head_on_collision_conserves_momentum_and_obeys_restitution(e1, e2, m2):
    """two circles, no friction: total momentum is conserved to rounding, and the relative normal velocity after the
    collision is -e1*e2 times the one before (Chipmunk multiplies the shapes' elasticities)"""
    sp = pymunk.Space()
    a = _circle(sp, 1.0, 5.0, (0, 0), (40, 0), e1)
    b = _circle(sp, m2, 5.0, (14, 0), (-10, 0), e2)
    p0, _ = _momentum([a, b])
    vrel0 = b.velocity[0] - a.velocity[0]
    for _ in range(60):
        sp.step(DT)
        p, _ = _momentum([a, b])
        assert np.allclose(p, p0, atol=1e-9 * (1 + np.abs(p0).max()))
    vrel1 = b.velocity[0] - a.velocity[0]
    assert vrel1 > -1e-9  # separating (or stuck together for e = 0)
    assert vrel1 == pytest.approx(-e1 * e2 * vrel0, rel=1e-6, abs=1e-6)


def test_oblique_collision_conserves_angular_momentum_and_stays_in_the_friction_cone():
    """a spinning circle grazing another.  Without friction linear AND angular momentum (about the origin) are conserved
    to rounding.  With friction the tangential impulse never exceeds u1*u2 times the normal impulse, and angular
    momentum changes only by what the penetration allows: Chipmunk applies the equal and opposite impulses at the two
    shapes' OWN surface points (cpArbiterPreStep: r1 = p1 - a.p, r2 = p2 - b.p), which lie a penetration depth apart."""
    for u1, u2 in ((0.0, 0.0), (0.8, 0.5)):
        sp = pymunk.Space()
        a = _circle(sp, 2.0, 5.0, (0, 0), (60, 0), 0.6, u1, w=3.0)
        b = _circle(sp, 1.0, 5.0, (13, 6), (0, 0), 0.6, u2)
        p0, L0 = _momentum([a, b])
        vb0 = np.array(b.velocity)
        for _ in range(40):
            sp.step(DT)
        p1, L1 = _momentum([a, b])
        assert np.allclose(p1, p0, atol=1e-9 * np.abs(p0).max())
        # impulse received by b, split along the contact normal at impact (centre line when the circles touch)
        J = b.mass * (np.array(b.velocity) - vb0)
        assert np.linalg.norm(J) > 1.0  # they did collide
        ax = 13 - math.sqrt(100 - 36)   # a moved along x until |b - a| = 10
        n = np.array([13 - ax, 6.0]) / 10.0
        jn, jt = J @ n, J @ np.array([-n[1], n[0]])
        # (the normal moves by up to the per-step travel / radius while the shapes overlap: 0.6 / 10)
        assert jn > 0 and abs(jt) <= u1 * u2 * jn * 1.02 + 0.07 * jn * (u1 * u2 == 0) + 1e-9
        if u1 * u2 == 0:
            # normal impulses pass through the centres: the spins are untouched (to rounding)
            assert a.angular_velocity == pytest.approx(3.0, abs=1e-12) and abs(b.angular_velocity) < 1e-12
        else:
            # the friction impulse j_t acts at the two surfaces, r = 5 from either centre: equal and opposite forces on
            # opposite sides give both circles the SAME torque impulse r x j_t (sliding friction spins both the same way)
            ta, tb = a.moment * (a.angular_velocity - 3.0), b.moment * b.angular_velocity
            assert ta == pytest.approx(tb, rel=1e-9) and 0.5 * 5.0 * abs(jt) < abs(tb) < 2.0 * 5.0 * abs(jt)  # (the normal turns while they touch)
        # (the orbital part of the angular momentum shifts by m * dx x v with the position correction dx of the overlap:
        # Chipmunk corrects penetration through a bias velocity that moves bodies without touching their momenta)
        assert abs(L1 - L0) <= (abs(jt) + 0.02 * jn) * 60 * DT * 2 + 1.0, (L0, L1, jt)


def test_penetration_settles_onto_collision_slop_at_the_bias_rate():
    """two overlapping circles at rest: the overlap beyond collision_slop decays by collision_bias^dt per step (the
    position correction of cpArbiterPreStep: bias = -(1 - collision_bias^dt) * min(0, dist + slop) / dt) and settles ON
    the slop, never separating the shapes"""
    sp = pymunk.Space()
    a = _circle(sp, 1.0, 5.0, (0, 0), e=0.0)
    b = _circle(sp, 1.0, 5.0, (9.0, 0), e=0.0)  # overlap 1.0, slop 0.1
    rate = sp.collision_bias ** DT
    excess = []
    for _ in range(400):
        sp.step(DT)
        excess.append(10.0 - (b.position[0] - a.position[0]) - sp.collision_slop)
    excess = np.array(excess)
    assert (excess > -1e-9).all()                       # never pushed beyond the slop
    assert (np.diff(excess) <= 1e-12).all()              # monotone
    ratios = excess[11:200] / excess[10:199]
    assert np.allclose(ratios, rate, rtol=2e-3), (ratios.min(), ratios.max(), rate)
    assert excess[-1] < 0.9 * rate ** 380 * 1.5          # and it keeps decaying at that rate
    assert abs(a.velocity[0]) < 1e-9 and abs(b.velocity[0]) < 1e-9  # the correction is a bias velocity: no real velocity gained


def test_pivot_joint_error_decays_by_error_bias():
    """a PivotJoint between two free bodies whose anchors are apart: the anchor distance shrinks by error_bias^dt per
    step (Robot.py:59 sets error_bias = 0.1 on the feet's joint), momentum is conserved"""
    sp = pymunk.Space()
    a = _circle(sp, 4000.0, 7.5, (100, 100))
    b = _circle(sp, 4000.0, 7.5, (103, 104))
    for s in sp.shapes:
        s.filter = pymunk.ShapeFilter(group=1)  # the feet of one robot do not collide through the filter
    j = PivotJoint(a, b, (0, 0), (0, 0))
    j.error_bias = 0.1
    sp.add(j)
    rate = 0.1 ** DT
    d = []
    for _ in range(120):
        sp.step(DT)
        d.append(math.hypot(b.position[0] - a.position[0], b.position[1] - a.position[1]))
        p, _ = _momentum([a, b])
        assert np.allclose(p, 0, atol=1e-6)
    d = np.array(d)
    assert np.allclose(d[6:100] / d[5:99], rate, rtol=1e-3), (d[:8], rate)
    com = (np.array(a.position) + np.array(b.position)) / 2
    assert np.allclose(com, [101.5, 102.0], atol=1e-9)  # the centre of mass stays


def test_rotary_limit_joint_pulls_the_angles_together():
    """RotaryLimitJoint(a, b, 0, 0) (Robot.py:60): the first correction removes (1 - error_bias^dt) of the angle difference
    per step; a limit joint only ever pushes towards its range (its accumulated impulse is clamped to one sign), so
    without the feet's rotational friction the difference then closes at that constant rate; angular momentum is
    conserved and the joint never drives the bodies apart"""
    sp = pymunk.Space()
    a = _circle(sp, 4000.0, 7.5, (0, 0))
    b = _circle(sp, 4000.0, 7.5, (0, 0))
    for s in sp.shapes:
        s.filter = pymunk.ShapeFilter(group=1)
    b.angle = 0.3
    j = RotaryLimitJoint(a, b, 0, 0)
    sp.add(j)
    rate = j.error_bias ** DT
    d = []
    for _ in range(100):
        sp.step(DT)
        d.append(b.angle - a.angle)
        assert a.moment * a.angular_velocity + b.moment * b.angular_velocity == pytest.approx(0.0, abs=1e-6)
    d = np.array(d)
    assert d[0] == pytest.approx(0.3) and d[1] == pytest.approx(0.3 * rate, rel=1e-9)  # positions integrate before the solve
    assert (np.diff(d[:15]) < 0).all() and np.allclose(np.diff(d[1:15]), d[1] - d[0], rtol=1e-6)
    assert np.abs(d[40:]).max() < 1e-3 and np.abs(d).max() <= 0.3 + 1e-12
    assert a.angle + b.angle == pytest.approx(0.3, abs=1e-9)


def test_resting_contact_is_a_fixed_point():
    """two circles touching within the slop, at rest: a step changes nothing (no drift, no jitter) -- what lets the
    kernels skip quiet pairs"""
    sp = pymunk.Space()
    a = _circle(sp, 1.0, 5.0, (0, 0), e=0.5, u=0.5)
    b = _circle(sp, 1.0, 5.0, (9.95, 0), e=0.5, u=0.5)  # overlap 0.05 < slop
    for _ in range(50):
        sp.step(DT)
    assert tuple(a.position) == (0.0, 0.0) and tuple(b.position) == (9.95, 0.0)
    assert tuple(a.velocity) == (0.0, 0.0) and tuple(b.velocity) == (0.0, 0.0)


def test_circle_against_static_segment_reflects():
    """a ball against a static wall segment (field posts / buildings are static shapes): the normal velocity is reflected
    with e1*e2, the tangential one kept without friction"""
    sp = pymunk.Space()
    wall = pymunk.Segment(sp.static_body, (-50, 0), (50, 0), 2.0)
    wall.elasticity, wall.friction = 0.95, 0.0
    sp.add(wall)
    ball = _circle(sp, 10.0, 5.0, (0, 12), (30, -80), 0.98)
    for _ in range(40):
        sp.step(DT)
    assert ball.velocity[1] == pytest.approx(80 * 0.95 * 0.98, rel=1e-6)
    assert ball.velocity[0] == pytest.approx(30, rel=1e-9)
