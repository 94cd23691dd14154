"""pymunk 5.x API stand-in: a CPU restatement of the Chipmunk2D 7.0.x subset DynEnv uses.

TEST INFRASTRUCTURE ONLY (oracle).  The real dependency (pymunk<6, unpinned in the
reference's requirements.txt:3 / setup.py:24) is not installable in this image, so the
rigid-body arithmetic is restated here from the published Chipmunk2D 7.0.x algorithm
(cpSpaceStep.c, cpArbiter.c, cpCollision.c, cpPivotJoint.c, cpRotaryLimitJoint.c, cpBody.c).
PARITY UNPINNED at this boundary: there is no pymunk wheel and the reference ships no
golden vectors, so nothing here can be checked against the true library.

With this package on sys.path the reference's own env files run unmodified
(oracle/ref_loader.py); every reference call site into pymunk is covered:
Space()/gravity/add/remove/step/add_collision_handler/point_query, Body (position, angle,
velocity, angular_velocity, velocity_func, mass, apply_force_at_world_point,
update_velocity), Circle/Segment/Poly, moment_for_*, PivotJoint, RotaryLimitJoint,
ShapeFilter and the mutable 5.x Vec2d.

Deterministic choices where Chipmunk's own order depends on its BB-tree internals:
arbiters are processed in ascending (shape index a, shape index b) order and for
same-type pairs shape `a` is the earlier-added one.
"""
import math

inf = float("inf")

# --------------------------------------------------------------------------- Vec2d


class Vec2d(object):
    """Mutable 2-D vector with the pymunk 5.x semantics the reference relies on."""
    __slots__ = ("x", "y")

    def __init__(self, x_or_pair=None, y=None):
        if x_or_pair is None:
            self.x = 0.0
            self.y = 0.0
        elif y is None:
            self.x = x_or_pair[0]
            self.y = x_or_pair[1]
        else:
            self.x = x_or_pair
            self.y = y

    def __len__(self):
        return 2

    def __getitem__(self, i):
        if i == 0:
            return self.x
        if i == 1:
            return self.y
        raise IndexError(i)

    def __setitem__(self, i, v):
        if i == 0:
            self.x = v
        elif i == 1:
            self.y = v
        else:
            raise IndexError(i)

    def __iter__(self):
        yield self.x
        yield self.y

    def __repr__(self):
        return "Vec2d(%s, %s)" % (self.x, self.y)

    def __eq__(self, o):
        if hasattr(o, "__getitem__") and len(o) == 2:
            return self.x == o[0] and self.y == o[1]
        return False

    def __ne__(self, o):
        if hasattr(o, "__getitem__") and len(o) == 2:
            return self.x != o[0] or self.y != o[1]
        return True

    __hash__ = None

    def __copy__(self):
        return Vec2d(self.x, self.y)

    def __deepcopy__(self, memo):
        return Vec2d(self.x, self.y)

    def __add__(self, o):
        return Vec2d(self.x + o[0], self.y + o[1])

    __radd__ = __add__

    def __iadd__(self, o):
        self.x += o[0]
        self.y += o[1]
        return self

    def __sub__(self, o):
        return Vec2d(self.x - o[0], self.y - o[1])

    def __rsub__(self, o):
        return Vec2d(o[0] - self.x, o[1] - self.y)

    def __isub__(self, o):
        self.x -= o[0]
        self.y -= o[1]
        return self

    def __mul__(self, o):
        if isinstance(o, (Vec2d, tuple, list)):
            return Vec2d(self.x * o[0], self.y * o[1])
        return Vec2d(self.x * o, self.y * o)

    __rmul__ = __mul__

    def __imul__(self, o):
        self.x *= o
        self.y *= o
        return self

    def __truediv__(self, o):
        return Vec2d(self.x / o, self.y / o)

    def __neg__(self):
        return Vec2d(-self.x, -self.y)

    def __pos__(self):
        return Vec2d(self.x, self.y)

    def __abs__(self):
        return Vec2d(abs(self.x), abs(self.y))

    def get_length_sqrd(self):
        return self.x * self.x + self.y * self.y

    def get_length(self):
        return math.sqrt(self.x * self.x + self.y * self.y)

    length = property(get_length)

    def rotate(self, angle_radians):
        c = math.cos(angle_radians)
        s = math.sin(angle_radians)
        x = self.x * c - self.y * s
        y = self.x * s + self.y * c
        self.x = x
        self.y = y

    def rotated(self, angle_radians):
        c = math.cos(angle_radians)
        s = math.sin(angle_radians)
        return Vec2d(self.x * c - self.y * s, self.x * s + self.y * c)

    def get_angle(self):
        if self.get_length_sqrd() == 0:
            return 0
        return math.atan2(self.y, self.x)

    angle = property(get_angle)

    def normalized(self):
        l = self.length
        if l != 0:
            return self / l
        return Vec2d(self)

    def perpendicular(self):
        return Vec2d(-self.y, self.x)

    def dot(self, o):
        return float(self.x * o[0] + self.y * o[1])

    def cross(self, o):
        return self.x * o[1] - self.y * o[0]

    def get_distance(self, o):
        return math.sqrt((self.x - o[0]) ** 2 + (self.y - o[1]) ** 2)

    @staticmethod
    def zero():
        return Vec2d(0, 0)

    @staticmethod
    def unit():
        return Vec2d(0, 1)


# --------------------------------------------------------------------------- moments


def moment_for_circle(mass, inner_radius, outer_radius, offset=(0, 0)):
    # cpMomentForCircle
    return mass * (0.5 * (inner_radius * inner_radius + outer_radius * outer_radius)
                   + (offset[0] * offset[0] + offset[1] * offset[1]))


def moment_for_segment(mass, a, b, radius):
    # cpMomentForSegment
    ox = 0.5 * (a[0] + b[0])
    oy = 0.5 * (a[1] + b[1])
    length = math.hypot(b[0] - a[0], b[1] - a[1]) + 2.0 * radius
    return mass * ((length * length + 4.0 * radius * radius) / 12.0 + (ox * ox + oy * oy))


def moment_for_poly(mass, vertices, offset=(0, 0), radius=0):
    # cpMomentForPoly
    n = len(vertices)
    if n == 2:
        return moment_for_segment(mass, vertices[0], vertices[1], 0.0)
    sum1 = 0.0
    sum2 = 0.0
    for i in range(n):
        v1x = vertices[i][0] + offset[0]
        v1y = vertices[i][1] + offset[1]
        v2x = vertices[(i + 1) % n][0] + offset[0]
        v2y = vertices[(i + 1) % n][1] + offset[1]
        a = v2x * v1y - v2y * v1x
        b = (v1x * v1x + v1y * v1y) + (v1x * v2x + v1y * v2y) + (v2x * v2x + v2y * v2y)
        sum1 += a * b
        sum2 += a
    return (mass * sum1) / (6.0 * sum2)


def moment_for_box(mass, size):
    return mass * (size[0] * size[0] + size[1] * size[1]) / 12.0


# --------------------------------------------------------------------------- bodies


class Body(object):
    DYNAMIC = 0
    KINEMATIC = 1
    STATIC = 2

    def __init__(self, mass=0, moment=0, body_type=DYNAMIC):
        self.body_type = body_type
        if body_type == Body.DYNAMIC:
            self._m = float(mass)
            self._i = float(moment)
        else:
            self._m = inf
            self._i = inf
        self._update_inv()
        self._px = 0.0
        self._py = 0.0
        self._a = 0.0
        self._vx = 0.0
        self._vy = 0.0
        self._w = 0.0
        self._fx = 0.0
        self._fy = 0.0
        self._t = 0.0
        self._vbx = 0.0
        self._vby = 0.0
        self._wb = 0.0
        self.velocity_func = None  # None -> default cpBodyUpdateVelocity
        self.shapes = []
        self.space = None

    def _update_inv(self):
        self._m_inv = 0.0 if self._m == inf else (inf if self._m == 0 else 1.0 / self._m)
        self._i_inv = 0.0 if self._i == inf else (inf if self._i == 0 else 1.0 / self._i)

    # getters return copies (pymunk 5: `body.position.x = ..` is a no-op on the body)
    def _get_position(self):
        return Vec2d(self._px, self._py)

    def _set_position(self, p):
        self._px = float(p[0])
        self._py = float(p[1])
        if self.body_type == Body.STATIC:
            for s in self.shapes:
                s.cache_bb()

    position = property(_get_position, _set_position)

    def _get_angle(self):
        return self._a

    def _set_angle(self, a):
        self._a = float(a)
        if self.body_type == Body.STATIC:
            for s in self.shapes:
                s.cache_bb()

    angle = property(_get_angle, _set_angle)

    def _get_velocity(self):
        return Vec2d(self._vx, self._vy)

    def _set_velocity(self, v):
        self._vx = float(v[0])
        self._vy = float(v[1])

    velocity = property(_get_velocity, _set_velocity)

    def _get_w(self):
        return self._w

    def _set_w(self, w):
        self._w = float(w)

    angular_velocity = property(_get_w, _set_w)

    def _get_mass(self):
        return self._m

    def _set_mass(self, m):
        self._m = float(m)
        self._update_inv()

    mass = property(_get_mass, _set_mass)

    def _get_moment(self):
        return self._i

    def _set_moment(self, i):
        self._i = float(i)
        self._update_inv()

    moment = property(_get_moment, _set_moment)

    def _get_force(self):
        return Vec2d(self._fx, self._fy)

    def _set_force(self, f):
        self._fx = float(f[0])
        self._fy = float(f[1])

    force = property(_get_force, _set_force)

    def _get_torque(self):
        return self._t

    def _set_torque(self, t):
        self._t = float(t)

    torque = property(_get_torque, _set_torque)

    @property
    def rotation_vector(self):
        return Vec2d(math.cos(self._a), math.sin(self._a))

    def apply_force_at_world_point(self, force, point):
        # cpBodyApplyForceAtWorldPoint (centre of gravity is the body origin everywhere here)
        fx = force[0]
        fy = force[1]
        self._fx += fx
        self._fy += fy
        rx = point[0] - self._px
        ry = point[1] - self._py
        self._t += rx * fy - ry * fx

    def apply_impulse_at_world_point(self, impulse, point):
        rx = point[0] - self._px
        ry = point[1] - self._py
        self._vx += impulse[0] * self._m_inv
        self._vy += impulse[1] * self._m_inv
        self._w += self._i_inv * (rx * impulse[1] - ry * impulse[0])

    def world_to_local(self, v):
        c = math.cos(self._a)
        s = math.sin(self._a)
        dx = v[0] - self._px
        dy = v[1] - self._py
        return Vec2d(dx * c + dy * s, -dx * s + dy * c)

    def local_to_world(self, v):
        c = math.cos(self._a)
        s = math.sin(self._a)
        return Vec2d(self._px + v[0] * c - v[1] * s, self._py + v[0] * s + v[1] * c)

    @staticmethod
    def update_velocity(body, gravity, damping, dt):
        # cpBodyUpdateVelocity
        if body.body_type == Body.KINEMATIC:
            return
        body._vx = body._vx * damping + (gravity[0] + body._fx * body._m_inv) * dt
        body._vy = body._vy * damping + (gravity[1] + body._fy * body._m_inv) * dt
        body._w = body._w * damping + body._t * body._i_inv * dt
        body._fx = 0.0
        body._fy = 0.0
        body._t = 0.0

    @staticmethod
    def update_position(body, dt):
        # cpBodyUpdatePosition
        body._px += (body._vx + body._vbx) * dt
        body._py += (body._vy + body._vby) * dt
        body._a += (body._w + body._wb) * dt
        body._vbx = 0.0
        body._vby = 0.0
        body._wb = 0.0


# --------------------------------------------------------------------------- shapes


class ShapeFilter(object):
    ALL_MASKS = 0xFFFFFFFF
    ALL_CATEGORIES = 0xFFFFFFFF

    def __init__(self, group=0, categories=0xFFFFFFFF, mask=0xFFFFFFFF):
        self.group = group
        self.categories = categories
        self.mask = mask

    def rejects(self, o):
        # cpShapeFilterReject
        return ((self.group != 0 and self.group == o.group)
                or (self.categories & o.mask) == 0
                or (o.categories & self.mask) == 0)


class PointQueryInfo(object):
    def __init__(self, shape, point, distance, gradient):
        self.shape = shape
        self.point = point
        self.distance = distance
        self.gradient = gradient


class Shape(object):
    TYPE = -1

    def __init__(self, body):
        self.body = body
        self.elasticity = 0.0
        self.friction = 0.0
        self.collision_type = 0
        self.sensor = False
        self.filter = ShapeFilter()
        self.color = None
        self.surface_velocity = Vec2d(0, 0)
        self.index = -1  # insertion index in the space
        self.bb = (0.0, 0.0, 0.0, 0.0)  # l, b, r, t
        if body is not None:
            body.shapes.append(self)

    def cache_bb(self):
        raise NotImplementedError


class Circle(Shape):
    TYPE = 0

    def __init__(self, body, radius, offset=(0, 0)):
        Shape.__init__(self, body)
        self.radius = float(radius)
        self.offset = Vec2d(offset)
        self.tc = (0.0, 0.0)
        self.cache_bb()

    def cache_bb(self):
        b = self.body
        c = math.cos(b._a)
        s = math.sin(b._a)
        x = b._px + self.offset.x * c - self.offset.y * s
        y = b._py + self.offset.x * s + self.offset.y * c
        self.tc = (x, y)
        r = self.radius
        self.bb = (x - r, y - r, x + r, y + r)

    def point_query(self, p):
        dx = p[0] - self.tc[0]
        dy = p[1] - self.tc[1]
        d = math.hypot(dx, dy)
        return d - self.radius


class Segment(Shape):
    TYPE = 1

    def __init__(self, body, a, b, radius):
        Shape.__init__(self, body)
        self.a = Vec2d(a)
        self.b = Vec2d(b)
        self.radius = float(radius)
        dx = self.b.x - self.a.x
        dy = self.b.y - self.a.y
        l = math.hypot(dx, dy)
        # seg->n = rperp(normalize(b - a))
        self.n = (dy / l, -dx / l)
        self.ta = (0.0, 0.0)
        self.tb = (0.0, 0.0)
        self.tn = (0.0, 0.0)
        self.cache_bb()

    def cache_bb(self):
        bd = self.body
        c = math.cos(bd._a)
        s = math.sin(bd._a)
        self.ta = (bd._px + self.a.x * c - self.a.y * s, bd._py + self.a.x * s + self.a.y * c)
        self.tb = (bd._px + self.b.x * c - self.b.y * s, bd._py + self.b.x * s + self.b.y * c)
        self.tn = (self.n[0] * c - self.n[1] * s, self.n[0] * s + self.n[1] * c)
        r = self.radius
        self.bb = (min(self.ta[0], self.tb[0]) - r, min(self.ta[1], self.tb[1]) - r,
                   max(self.ta[0], self.tb[0]) + r, max(self.ta[1], self.tb[1]) + r)

    def point_query(self, p):
        cx, cy = _closest_on_segment(p[0], p[1], self.ta, self.tb)
        return math.hypot(p[0] - cx, p[1] - cy) - self.radius


def _convex_hull_ccw(pts):
    # Andrew monotone chain; returns CCW hull (cpConvexHull yields a CCW loop too; the start
    # vertex differs, which only relabels feature ids).
    p = sorted(set((float(x), float(y)) for x, y in pts))
    if len(p) <= 2:
        return p

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower = []
    for q in p:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], q) <= 0:
            lower.pop()
        lower.append(q)
    upper = []
    for q in reversed(p):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], q) <= 0:
            upper.pop()
        upper.append(q)
    return lower[:-1] + upper[:-1]


class Poly(Shape):
    TYPE = 2

    def __init__(self, body, vertices, transform=None, radius=0):
        Shape.__init__(self, body)
        self.radius = float(radius)
        self.local = _convex_hull_ccw([(v[0], v[1]) for v in vertices])
        n = len(self.local)
        # plane i: edge from vert[i-1] to vert[i], outward normal rperp(normalize(b-a)), v0 = vert[i]
        self.local_n = []
        for i in range(n):
            a = self.local[(i - 1 + n) % n]
            b = self.local[i]
            dx = b[0] - a[0]
            dy = b[1] - a[1]
            l = math.hypot(dx, dy)
            self.local_n.append((dy / l, -dx / l))
        self.tv = [(0.0, 0.0)] * n
        self.tn = [(0.0, 0.0)] * n
        self.cache_bb()

    def get_vertices(self):
        return [Vec2d(v) for v in self.local]

    def cache_bb(self):
        bd = self.body
        c = math.cos(bd._a)
        s = math.sin(bd._a)
        self.tv = [(bd._px + x * c - y * s, bd._py + x * s + y * c) for x, y in self.local]
        self.tn = [(x * c - y * s, x * s + y * c) for x, y in self.local_n]
        r = self.radius
        xs = [v[0] for v in self.tv]
        ys = [v[1] for v in self.tv]
        self.bb = (min(xs) - r, min(ys) - r, max(xs) + r, max(ys) + r)

    def point_query(self, p):
        n = len(self.tv)
        inside = True
        best = inf
        for i in range(n):
            a = self.tv[(i - 1 + n) % n]
            b = self.tv[i]
            if (p[0] - b[0]) * self.tn[i][0] + (p[1] - b[1]) * self.tn[i][1] > 0:
                inside = False
            cx, cy = _closest_on_segment(p[0], p[1], a, b)
            best = min(best, math.hypot(p[0] - cx, p[1] - cy))
        return (-best if inside else best) - self.radius


def _closest_on_segment(px, py, a, b):
    dx = b[0] - a[0]
    dy = b[1] - a[1]
    t = ((px - a[0]) * dx + (py - a[1]) * dy) / (dx * dx + dy * dy)
    t = 0.0 if t < 0.0 else (1.0 if t > 1.0 else t)
    return a[0] + dx * t, a[1] + dy * t


# --------------------------------------------------------------------------- narrow phase


class Contact(object):
    __slots__ = ("r1x", "r1y", "r2x", "r2y", "hash", "jnAcc", "jtAcc", "jBias", "nMass",
                 "tMass", "bias", "bounce")

    def __init__(self, p1, p2, h):
        # absolute points first; converted to body-relative offsets in Arbiter.update
        self.r1x, self.r1y = p1
        self.r2x, self.r2y = p2
        self.hash = h
        self.jnAcc = 0.0
        self.jtAcc = 0.0
        self.jBias = 0.0
        self.nMass = 0.0
        self.tMass = 0.0
        self.bias = 0.0
        self.bounce = 0.0


def _circle_circle(p1, p2, r1, r2):
    # CircleToCircleQuery
    mind = r1 + r2
    dx = p2[0] - p1[0]
    dy = p2[1] - p1[1]
    dsq = dx * dx + dy * dy
    if dsq < mind * mind:
        d = math.sqrt(dsq)
        n = (dx / d, dy / d) if d else (1.0, 0.0)
        return n, [Contact((p1[0] + n[0] * r1, p1[1] + n[1] * r1),
                           (p2[0] - n[0] * r2, p2[1] - n[1] * r2), 0)]
    return None, []


def _circle_segment(circ, seg):
    # CircleToSegment (end-cap tangents are zero here => never rejected)
    a = seg.ta
    b = seg.tb
    c = circ.tc
    sdx = b[0] - a[0]
    sdy = b[1] - a[1]
    t = (sdx * (c[0] - a[0]) + sdy * (c[1] - a[1])) / (sdx * sdx + sdy * sdy)
    t = 0.0 if t < 0.0 else (1.0 if t > 1.0 else t)
    clx = a[0] + sdx * t
    cly = a[1] + sdy * t
    mind = circ.radius + seg.radius
    dx = clx - c[0]
    dy = cly - c[1]
    dsq = dx * dx + dy * dy
    if dsq < mind * mind:
        d = math.sqrt(dsq)
        n = (dx / d, dy / d) if d else seg.tn
        return n, [Contact((c[0] + n[0] * circ.radius, c[1] + n[1] * circ.radius),
                           (clx - n[0] * seg.radius, cly - n[1] * seg.radius), 0)]
    return None, []


def _core_points(s):
    if s.TYPE == 1:
        return [s.ta, s.tb]
    if s.TYPE == 2:
        return list(s.tv)
    return [s.tc]


def _edge_axes(pts):
    """Outward unit normals of a CCW polygon (a segment contributes its two sides)."""
    n = len(pts)
    out = []
    if n == 2:
        dx = pts[1][0] - pts[0][0]
        dy = pts[1][1] - pts[0][1]
        l = math.hypot(dx, dy)
        out.append((dy / l, -dx / l))
        out.append((-dy / l, dx / l))
        return out
    for i in range(n):
        a = pts[i]
        b = pts[(i + 1) % n]
        dx = b[0] - a[0]
        dy = b[1] - a[1]
        l = math.hypot(dx, dy)
        out.append((dy / l, -dx / l))
    return out


def _closest_points_convex(A, B):
    """(n, d) of GJK/EPA ClosestPoints for two convex cores A, B (point lists, CCW).

    n points from A towards B.  Separated: d = core distance, n = unit(pb - pa).
    Overlapping: minimum-penetration axis among both shapes' edge normals, d <= 0.
    """
    best_d = -inf
    best_n = None
    for ax in _edge_axes(A):  # candidate n = +axis
        d = min(ax[0] * q[0] + ax[1] * q[1] for q in B) - max(ax[0] * q[0] + ax[1] * q[1] for q in A)
        if d > best_d:
            best_d = d
            best_n = ax
    for ax in _edge_axes(B):  # candidate n = -axis
        d = min(ax[0] * q[0] + ax[1] * q[1] for q in A) - max(ax[0] * q[0] + ax[1] * q[1] for q in B)
        if d > best_d:
            best_d = d
            best_n = (-ax[0], -ax[1])
    if best_d <= 0.0 and _cores_overlap(A, B):
        return best_n, best_d
    # separated: closest feature pair (vertex-edge or vertex-vertex)
    best = inf
    bn = best_n

    def scan(P, Q, sign):
        nonlocal best, bn
        m = len(Q)
        edges = [(Q[i], Q[(i + 1) % m]) for i in range(m if m > 2 else 1)] if m > 1 else []
        for p in P:
            if not edges:
                cands = [Q[0]]
            else:
                cands = [_closest_on_segment(p[0], p[1], e[0], e[1]) for e in edges]
            for c in cands:
                dx = c[0] - p[0]
                dy = c[1] - p[1]
                d = math.hypot(dx, dy)
                if d < best:
                    best = d
                    if d > 0:
                        bn = (sign * dx / d, sign * dy / d)

    scan(A, B, 1.0)
    scan(B, A, -1.0)
    return bn, best


def _cores_overlap(A, B):
    """True when the convex cores intersect (SAT incl. vertex axes is overkill here: for the
    degenerate 2-gon case edge normals alone can miss a separation, so test segments directly)."""
    if len(A) == 2 and len(B) == 2:
        return _segments_intersect(A[0], A[1], B[0], B[1])
    if len(A) == 2 or len(B) == 2:
        S, P = (A, B) if len(A) == 2 else (B, A)
        if _point_in_poly(S[0], P) or _point_in_poly(S[1], P):
            return True
        m = len(P)
        return any(_segments_intersect(S[0], S[1], P[i], P[(i + 1) % m]) for i in range(m))
    return True  # poly-poly: edge-normal SAT is exact


def _point_in_poly(p, P):
    m = len(P)
    for i in range(m):
        a = P[i]
        b = P[(i + 1) % m]
        if (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) < 0:
            return False
    return True


def _segments_intersect(p1, p2, p3, p4):
    def orient(a, b, c):
        return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])

    d1 = orient(p3, p4, p1)
    d2 = orient(p3, p4, p2)
    d3 = orient(p1, p2, p3)
    d4 = orient(p1, p2, p4)
    return (d1 * d2 < 0) and (d3 * d4 < 0)


def _support_edge(s, n):
    """SupportEdgeForSegment / SupportEdgeForPoly: ((pa, ha), (pb, hb), r)."""
    if s.TYPE == 1:
        if s.tn[0] * n[0] + s.tn[1] * n[1] > 0.0:
            return (s.ta, (s.index, 0)), (s.tb, (s.index, 1)), s.radius
        return (s.tb, (s.index, 1)), (s.ta, (s.index, 0)), s.radius
    cnt = len(s.tv)
    i1 = max(range(cnt), key=lambda i: (s.tv[i][0] * n[0] + s.tv[i][1] * n[1], -i))
    i0 = (i1 - 1 + cnt) % cnt
    i2 = (i1 + 1) % cnt
    if (n[0] * s.tn[i1][0] + n[1] * s.tn[i1][1]) > (n[0] * s.tn[i2][0] + n[1] * s.tn[i2][1]):
        return (s.tv[i0], (s.index, i0)), (s.tv[i1], (s.index, i1)), s.radius
    return (s.tv[i1], (s.index, i1)), (s.tv[i2], (s.index, i2)), s.radius


_FLT_MIN = 2.2250738585072014e-308


def _clamp01(t):
    return 0.0 if t < 0.0 else (1.0 if t > 1.0 else t)


def _lerp(a, b, t):
    return (a[0] * (1.0 - t) + b[0] * t, a[1] * (1.0 - t) + b[1] * t)


def _contact_points(e1, e2, n, d):
    # ContactPoints (cpCollision.c)
    (e1a, h1a), (e1b, h1b), r1 = e1
    (e2a, h2a), (e2b, h2b), r2 = e2
    out = []
    if d <= r1 + r2:
        d_e1_a = e1a[0] * n[1] - e1a[1] * n[0]
        d_e1_b = e1b[0] * n[1] - e1b[1] * n[0]
        d_e2_a = e2a[0] * n[1] - e2a[1] * n[0]
        d_e2_b = e2b[0] * n[1] - e2b[1] * n[0]
        e1_denom = 1.0 / (d_e1_b - d_e1_a + _FLT_MIN)
        e2_denom = 1.0 / (d_e2_b - d_e2_a + _FLT_MIN)
        q1 = _lerp(e1a, e1b, _clamp01((d_e2_b - d_e1_a) * e1_denom))
        q2 = _lerp(e2a, e2b, _clamp01((d_e1_a - d_e2_a) * e2_denom))
        p1 = (q1[0] + n[0] * r1, q1[1] + n[1] * r1)
        p2 = (q2[0] - n[0] * r2, q2[1] - n[1] * r2)
        if (p2[0] - p1[0]) * n[0] + (p2[1] - p1[1]) * n[1] <= 0.0:
            out.append(Contact(p1, p2, (h1a, h2b)))
        q1 = _lerp(e1a, e1b, _clamp01((d_e2_a - d_e1_a) * e1_denom))
        q2 = _lerp(e2a, e2b, _clamp01((d_e1_b - d_e2_a) * e2_denom))
        p1 = (q1[0] + n[0] * r1, q1[1] + n[1] * r1)
        p2 = (q2[0] - n[0] * r2, q2[1] - n[1] * r2)
        if (p2[0] - p1[0]) * n[0] + (p2[1] - p1[1]) * n[1] <= 0.0:
            out.append(Contact(p1, p2, (h1b, h2a)))
    return out


def _circle_poly(circ, poly):
    # CircleToPoly: closest point of the polygon core to the circle centre
    c = circ.tc
    P = poly.tv
    m = len(P)
    if _point_in_poly(c, P):
        best = -inf
        bi = 0
        for i in range(m):  # plane i: v0 = P[i], outward normal tn[i]
            d = (c[0] - P[i][0]) * poly.tn[i][0] + (c[1] - P[i][1]) * poly.tn[i][1]
            if d > best:
                best = d
                bi = i
        n = (-poly.tn[bi][0], -poly.tn[bi][1])
        pb = _closest_on_segment(c[0], c[1], P[(bi - 1 + m) % m], P[bi])
        d = best
    else:
        d = inf
        pb = None
        for i in range(m):
            q = _closest_on_segment(c[0], c[1], P[(i - 1 + m) % m], P[i])
            dd = math.hypot(q[0] - c[0], q[1] - c[1])
            if dd < d:
                d = dd
                pb = q
        n = ((pb[0] - c[0]) / d, (pb[1] - c[1]) / d) if d > 0 else (1.0, 0.0)
    if d <= circ.radius + poly.radius:
        return n, [Contact((c[0] + n[0] * circ.radius, c[1] + n[1] * circ.radius),
                           (pb[0] - n[0] * poly.radius, pb[1] - n[1] * poly.radius), 0)]
    return None, []


def collide(a, b):
    """cpCollide: returns (a, b, n, contacts) with shapes ordered by primitive type."""
    if a.TYPE > b.TYPE:
        a, b = b, a
    ta, tb = a.TYPE, b.TYPE
    if ta == 0 and tb == 0:
        n, cs = _circle_circle(a.tc, b.tc, a.radius, b.radius)
    elif ta == 0 and tb == 1:
        n, cs = _circle_segment(a, b)
    elif ta == 0 and tb == 2:
        n, cs = _circle_poly(a, b)
    else:
        n, d = _closest_points_convex(_core_points(a), _core_points(b))
        cs = []
        if d - a.radius - b.radius <= 0.0:
            cs = _contact_points(_support_edge(a, n), _support_edge(b, (-n[0], -n[1])), n, d)
    return a, b, n, cs


# --------------------------------------------------------------------------- arbiters

_ST_FIRST, _ST_NORMAL, _ST_IGNORE, _ST_CACHED = 0, 1, 2, 3


class CollisionHandler(object):
    def __init__(self, type_a, type_b):
        self.type_a = type_a
        self.type_b = type_b
        self.begin = None
        self.pre_solve = None
        self.post_solve = None
        self.separate = None
        self.data = {}


class Arbiter(object):
    def __init__(self, a, b):
        self.a = a
        self.b = b
        self.contacts = []
        self.n = (0.0, 0.0)
        self.e = 0.0
        self.u = 0.0
        self.state = _ST_FIRST
        self.stamp = 0
        self.handler = None
        self.swapped = False

    @property
    def shapes(self):
        return (self.b, self.a) if self.swapped else (self.a, self.b)

    @property
    def is_first_contact(self):
        return self.state == _ST_FIRST

    @property
    def normal(self):
        return Vec2d(self.n) if not self.swapped else Vec2d(-self.n[0], -self.n[1])

    @property
    def total_impulse(self):
        sx = 0.0
        sy = 0.0
        for c in self.contacts:
            sx += self.n[0] * c.jnAcc - self.n[1] * c.jtAcc
            sy += self.n[0] * c.jtAcc + self.n[1] * c.jnAcc
        return Vec2d(-sx, -sy) if self.swapped else Vec2d(sx, sy)


def _k_scalar(a, b, r1x, r1y, r2x, r2y, nx, ny):
    r1cn = r1x * ny - r1y * nx
    r2cn = r2x * ny - r2y * nx
    return a._m_inv + a._i_inv * r1cn * r1cn + b._m_inv + b._i_inv * r2cn * r2cn


def _apply_impulses(a, b, r1x, r1y, r2x, r2y, jx, jy):
    a._vx -= jx * a._m_inv
    a._vy -= jy * a._m_inv
    a._w += a._i_inv * (r1x * -jy - r1y * -jx)
    b._vx += jx * b._m_inv
    b._vy += jy * b._m_inv
    b._w += b._i_inv * (r2x * jy - r2y * jx)


def _apply_bias_impulses(a, b, r1x, r1y, r2x, r2y, jx, jy):
    a._vbx -= jx * a._m_inv
    a._vby -= jy * a._m_inv
    a._wb += a._i_inv * (r1x * -jy - r1y * -jx)
    b._vbx += jx * b._m_inv
    b._vby += jy * b._m_inv
    b._wb += b._i_inv * (r2x * jy - r2y * jx)


# --------------------------------------------------------------------------- space


class Space(object):
    def __init__(self, threaded=False):
        self.gravity = (0.0, 0.0)
        self.iterations = 10
        self.damping = 1.0
        self.collision_slop = 0.1
        self.collision_bias = (1.0 - 0.1) ** 60.0
        self.collision_persistence = 3
        self.bodies = []  # dynamic bodies, insertion order
        self.static_bodies = []
        self.shapes = []  # all shapes, insertion order (index = feature/ordering id)
        self.constraints = []
        self.handlers = {}
        self.cached_arbiters = {}
        self.arbiters = []
        self.stamp = 0
        self.curr_dt = 0.0
        self._locked = False
        self._add_later = []
        self._remove_later = []
        self.static_body = Body(body_type=Body.STATIC)

    # -- membership
    def add(self, *objs):
        for o in objs:
            if isinstance(o, (list, tuple, set)):
                self.add(*o)
                continue
            if self._locked:
                self._add_later.append(o)
                continue
            if isinstance(o, Body):
                o.space = self
                (self.bodies if o.body_type == Body.DYNAMIC else self.static_bodies).append(o)
            elif isinstance(o, Shape):
                o.index = len(self.shapes)
                self.shapes.append(o)
                o.cache_bb()
            else:
                self.constraints.append(o)

    def remove(self, *objs):
        for o in objs:
            if isinstance(o, (list, tuple, set)):
                self.remove(*o)
                continue
            if self._locked:
                self._remove_later.append(o)
                continue
            if isinstance(o, Body):
                (self.bodies if o.body_type == Body.DYNAMIC else self.static_bodies).remove(o)
            elif isinstance(o, Shape):
                self.shapes.remove(o)
            else:
                # cpArrayDeleteObj: the last element takes the freed slot
                i = self.constraints.index(o)
                last = self.constraints.pop()
                if i < len(self.constraints):
                    self.constraints[i] = last

    def add_collision_handler(self, type_a, type_b):
        key = (type_a, type_b)
        if key in self.handlers:
            return self.handlers[key]
        if (type_b, type_a) in self.handlers:
            return self.handlers[(type_b, type_a)]
        h = CollisionHandler(type_a, type_b)
        self.handlers[key] = h
        return h

    def _lookup_handler(self, ta, tb):
        h = self.handlers.get((ta, tb))
        if h is None:
            h = self.handlers.get((tb, ta))
        return h

    def point_query(self, point, max_distance, shape_filter):
        out = []
        for s in self.shapes:
            if s.filter.rejects(shape_filter):
                continue
            d = s.point_query(point)
            if d < max_distance:
                out.append(PointQueryInfo(s, None, d, None))
        return out

    def debug_draw(self, options):
        pass

    # -- the step (cpSpaceStep)
    def step(self, dt):
        if dt == 0.0:
            return
        self.stamp += 1
        prev_dt = self.curr_dt
        self.curr_dt = dt
        for arb in self.arbiters:
            arb.state = _ST_NORMAL
        self.arbiters = []
        self._locked = True

        # integrate positions
        for b in self.bodies:
            Body.update_position(b, dt)

        # refresh world geometry of shapes on dynamic bodies, then collide
        for s in self.shapes:
            if s.body.body_type != Body.STATIC:
                s.cache_bb()
        shapes = self.shapes
        ns = len(shapes)
        for i in range(ns):
            si = shapes[i]
            for j in range(i + 1, ns):
                sj = shapes[j]
                if si.body.body_type == Body.STATIC and sj.body.body_type == Body.STATIC:
                    continue
                self._collide_shapes(si, sj)

        # separate callbacks / cache expiry (cpSpaceArbiterSetFilter)
        for key in list(self.cached_arbiters.keys()):
            arb = self.cached_arbiters[key]
            ticks = self.stamp - arb.stamp
            if ticks >= 1 and arb.state != _ST_CACHED:
                arb.state = _ST_CACHED
                h = arb.handler
                if h is not None and h.separate is not None:
                    h.separate(arb, self, h.data)
            if ticks >= self.collision_persistence:
                arb.contacts = []
                del self.cached_arbiters[key]

        # pre-step
        slop = self.collision_slop
        bias_coef = 1.0 - math.pow(self.collision_bias, dt)
        for arb in self.arbiters:
            a = arb.a.body
            b = arb.b.body
            nx, ny = arb.n
            bdx = b._px - a._px
            bdy = b._py - a._py
            for c in arb.contacts:
                c.nMass = 1.0 / _k_scalar(a, b, c.r1x, c.r1y, c.r2x, c.r2y, nx, ny)
                c.tMass = 1.0 / _k_scalar(a, b, c.r1x, c.r1y, c.r2x, c.r2y, -ny, nx)
                dist = (c.r2x - c.r1x + bdx) * nx + (c.r2y - c.r1y + bdy) * ny
                c.bias = -bias_coef * min(0.0, dist + slop) / dt
                c.jBias = 0.0
                vrx = (b._vx - c.r2y * b._w) - (a._vx - c.r1y * a._w)
                vry = (b._vy + c.r2x * b._w) - (a._vy + c.r1x * a._w)
                c.bounce = (vrx * nx + vry * ny) * arb.e
        for con in self.constraints:
            con.pre_step(dt)

        # integrate velocities
        damping = math.pow(self.damping, dt)
        g = Vec2d(self.gravity[0], self.gravity[1])
        for b in self.bodies:
            if b.velocity_func is None:
                Body.update_velocity(b, g, damping, dt)
            else:
                b.velocity_func(b, g, damping, dt)

        # warm start
        dt_coef = 0.0 if prev_dt == 0.0 else dt / prev_dt
        for arb in self.arbiters:
            if arb.state == _ST_FIRST:
                continue
            a = arb.a.body
            b = arb.b.body
            nx, ny = arb.n
            for c in arb.contacts:
                jx = nx * c.jnAcc - ny * c.jtAcc
                jy = nx * c.jtAcc + ny * c.jnAcc
                _apply_impulses(a, b, c.r1x, c.r1y, c.r2x, c.r2y, jx * dt_coef, jy * dt_coef)
        for con in self.constraints:
            con.apply_cached_impulse(dt_coef)

        # solver
        for _ in range(self.iterations):
            for arb in self.arbiters:
                self._arbiter_apply_impulse(arb)
            for con in self.constraints:
                con.apply_impulse(dt)

        # post-solve callbacks
        for arb in self.arbiters:
            h = arb.handler
            if h is not None and h.post_solve is not None:
                h.post_solve(arb, self, h.data)

        self._locked = False
        later_add, self._add_later = self._add_later, []
        later_rm, self._remove_later = self._remove_later, []
        for o in later_add:
            self.add(o)
        for o in later_rm:
            self.remove(o)

    def _collide_shapes(self, sa, sb):
        # QueryReject
        A = sa.bb
        B = sb.bb
        if not (A[0] <= B[2] and B[0] <= A[2] and A[1] <= B[3] and B[1] <= A[3]):
            return
        if sa.body is sb.body:
            return
        if sa.filter.rejects(sb.filter):
            return
        a, b, n, contacts = collide(sa, sb)
        if not contacts:
            return
        key = (a.index, b.index)
        arb = self.cached_arbiters.get(key)
        if arb is None:
            arb = Arbiter(a, b)
            self.cached_arbiters[key] = arb
        # cpArbiterUpdate
        arb.a = a
        arb.b = b
        for c in contacts:
            c.r1x -= a.body._px
            c.r1y -= a.body._py
            c.r2x -= b.body._px
            c.r2y -= b.body._py
            for old in arb.contacts:
                if c.hash == old.hash:
                    c.jnAcc = old.jnAcc
                    c.jtAcc = old.jtAcc
        arb.contacts = contacts
        arb.n = n
        arb.e = a.elasticity * b.elasticity
        arb.u = a.friction * b.friction
        h = self._lookup_handler(a.collision_type, b.collision_type)
        arb.handler = h
        arb.swapped = (h is not None and a.collision_type != h.type_a)
        if arb.state == _ST_CACHED:
            arb.state = _ST_FIRST
        if arb.state == _ST_FIRST and h is not None and h.begin is not None:
            if not h.begin(arb, self, h.data):
                arb.state = _ST_IGNORE
        ok = arb.state != _ST_IGNORE
        if ok and h is not None and h.pre_solve is not None:
            ok = bool(h.pre_solve(arb, self, h.data)) and arb.state != _ST_IGNORE
        if ok and not (a.sensor or b.sensor) and not (a.body._m == inf and b.body._m == inf):
            self.arbiters.append(arb)
        else:
            arb.contacts = []
            if arb.state != _ST_IGNORE:
                arb.state = _ST_NORMAL
        arb.stamp = self.stamp

    @staticmethod
    def _arbiter_apply_impulse(arb):
        # cpArbiterApplyImpulse (surface velocities are zero everywhere here)
        a = arb.a.body
        b = arb.b.body
        nx, ny = arb.n
        u = arb.u
        for c in arb.contacts:
            r1x, r1y, r2x, r2y = c.r1x, c.r1y, c.r2x, c.r2y
            vb1x = a._vbx - r1y * a._wb
            vb1y = a._vby + r1x * a._wb
            vb2x = b._vbx - r2y * b._wb
            vb2y = b._vby + r2x * b._wb
            vrx = (b._vx - r2y * b._w) - (a._vx - r1y * a._w)
            vry = (b._vy + r2x * b._w) - (a._vy + r1x * a._w)
            vbn = (vb2x - vb1x) * nx + (vb2y - vb1y) * ny
            vrn = vrx * nx + vry * ny
            vrt = -vrx * ny + vry * nx
            jbn = (c.bias - vbn) * c.nMass
            jbnOld = c.jBias
            c.jBias = max(jbnOld + jbn, 0.0)
            jn = -(c.bounce + vrn) * c.nMass
            jnOld = c.jnAcc
            c.jnAcc = max(jnOld + jn, 0.0)
            jtMax = u * c.jnAcc
            jt = -vrt * c.tMass
            jtOld = c.jtAcc
            c.jtAcc = min(max(jtOld + jt, -jtMax), jtMax)
            jb = c.jBias - jbnOld
            _apply_bias_impulses(a, b, r1x, r1y, r2x, r2y, nx * jb, ny * jb)
            dn = c.jnAcc - jnOld
            dtt = c.jtAcc - jtOld
            _apply_impulses(a, b, r1x, r1y, r2x, r2y, nx * dn - ny * dtt, nx * dtt + ny * dn)


from . import constraint, shape_filter, pygame_util  # noqa: E402,F401
from .constraint import PivotJoint, RotaryLimitJoint  # noqa: E402,F401

version = "5.7.0-shim"
chipmunk_version = "7.0.3-restated"
