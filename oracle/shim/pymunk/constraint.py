"""PivotJoint / RotaryLimitJoint restated from Chipmunk2D 7.0.x (cpPivotJoint.c,
cpRotaryLimitJoint.c).  Oracle/test infrastructure only; parity unpinned (see package doc)."""
import math

inf = float("inf")


def _bias_coef(error_bias, dt):
    return 1.0 - math.pow(error_bias, dt)


class Constraint(object):
    def __init__(self, a, b):
        self.a = a
        self.b = b
        self.max_force = inf
        self.error_bias = math.pow(1.0 - 0.1, 60.0)
        self.max_bias = inf
        self.collide_bodies = True


class PivotJoint(Constraint):
    def __init__(self, a, b, *args):
        Constraint.__init__(self, a, b)
        if len(args) == 1:
            pivot = args[0]
            self.anchor_a = tuple(a.world_to_local(pivot))
            self.anchor_b = tuple(b.world_to_local(pivot))
        else:
            self.anchor_a = (float(args[0][0]), float(args[0][1]))
            self.anchor_b = (float(args[1][0]), float(args[1][1]))
        self.jAcc = (0.0, 0.0)
        self.r1 = (0.0, 0.0)
        self.r2 = (0.0, 0.0)
        self.k = (0.0, 0.0, 0.0, 0.0)
        self.bias = (0.0, 0.0)

    def pre_step(self, dt):
        a, b = self.a, self.b
        ca, sa = math.cos(a._a), math.sin(a._a)
        cb, sb = math.cos(b._a), math.sin(b._a)
        r1 = (self.anchor_a[0] * ca - self.anchor_a[1] * sa, self.anchor_a[0] * sa + self.anchor_a[1] * ca)
        r2 = (self.anchor_b[0] * cb - self.anchor_b[1] * sb, self.anchor_b[0] * sb + self.anchor_b[1] * cb)
        self.r1, self.r2 = r1, r2
        # k_tensor, stored inverted
        m_sum = a._m_inv + b._m_inv
        k11 = m_sum + r1[1] * r1[1] * a._i_inv + r2[1] * r2[1] * b._i_inv
        k12 = -r1[0] * r1[1] * a._i_inv - r2[0] * r2[1] * b._i_inv
        k22 = m_sum + r1[0] * r1[0] * a._i_inv + r2[0] * r2[0] * b._i_inv
        det_inv = 1.0 / (k11 * k22 - k12 * k12)
        self.k = (k22 * det_inv, -k12 * det_inv, -k12 * det_inv, k11 * det_inv)
        dx = (b._px + r2[0]) - (a._px + r1[0])
        dy = (b._py + r2[1]) - (a._py + r1[1])
        c = -_bias_coef(self.error_bias, dt) / dt
        self.bias = (dx * c, dy * c)  # maxBias is infinite -> no clamp

    def apply_cached_impulse(self, dt_coef):
        self._apply(self.jAcc[0] * dt_coef, self.jAcc[1] * dt_coef)

    def _apply(self, jx, jy):
        a, b = self.a, self.b
        r1, r2 = self.r1, self.r2
        a._vx -= jx * a._m_inv
        a._vy -= jy * a._m_inv
        a._w += a._i_inv * (r1[0] * -jy - r1[1] * -jx)
        b._vx += jx * b._m_inv
        b._vy += jy * b._m_inv
        b._w += b._i_inv * (r2[0] * jy - r2[1] * jx)

    def apply_impulse(self, dt):
        a, b = self.a, self.b
        r1, r2 = self.r1, self.r2
        vrx = (b._vx - r2[1] * b._w) - (a._vx - r1[1] * a._w)
        vry = (b._vy + r2[0] * b._w) - (a._vy + r1[0] * a._w)
        ex = self.bias[0] - vrx
        ey = self.bias[1] - vry
        jx = self.k[0] * ex + self.k[1] * ey
        jy = self.k[2] * ex + self.k[3] * ey
        # maxForce infinite -> no clamp
        self.jAcc = (self.jAcc[0] + jx, self.jAcc[1] + jy)
        self._apply(jx, jy)


class RotaryLimitJoint(Constraint):
    def __init__(self, a, b, min, max):
        Constraint.__init__(self, a, b)
        self.min = float(min)
        self.max = float(max)
        self.jAcc = 0.0
        self.bias = 0.0
        self.iSum = 0.0

    def pre_step(self, dt):
        a, b = self.a, self.b
        dist = b._a - a._a
        pdist = 0.0
        if dist > self.max:
            pdist = self.max - dist
        elif dist < self.min:
            pdist = self.min - dist
        self.iSum = 1.0 / (a._i_inv + b._i_inv)
        self.bias = -_bias_coef(self.error_bias, dt) * pdist / dt
        if not self.bias:
            self.jAcc = 0.0

    def apply_cached_impulse(self, dt_coef):
        j = self.jAcc * dt_coef
        self.a._w -= j * self.a._i_inv
        self.b._w += j * self.b._i_inv

    def apply_impulse(self, dt):
        if not self.bias:
            return
        a, b = self.a, self.b
        wr = b._w - a._w
        j = -(self.bias + wr) * self.iSum
        jOld = self.jAcc
        if self.bias < 0.0:
            self.jAcc = max(jOld + j, 0.0)
        else:
            self.jAcc = min(jOld + j, 0.0)
        j = self.jAcc - jOld
        a._w -= j * a._i_inv
        b._w += j * b._i_inv
