"""Host-side value types and exceptions of the drop-in API.

Only what scene-building scripts touch: a 2-vector with the reference's operator
conventions (`pypybox2d/common.py:88-225`, conventions listed in SURVEY A.2) and the
exception types (`common.py:84-86`).  All per-step math lives on the device.
"""
import math
from numbers import Number

from .settings import EPSILON


class PhysicsError(Exception):
    pass


class LockedError(PhysicsError):
    pass


class EmptyFixtureError(PhysicsError):
    pass


class Vec2(object):
    """2-vector of Python floats; `+=` rebinds like the reference's (no in-place mutation)."""
    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=0.0):
        self.x = float(x)
        self.y = float(y)

    def __iter__(self):
        yield self.x
        yield self.y

    def __len__(self):
        return 2

    def __getitem__(self, i):
        if i == 0:
            return self.x
        if i == 1:
            return self.y
        raise IndexError("Index must be in (0,1)")

    def __setitem__(self, i, value):
        if i == 0:
            self.x = float(value)
        elif i == 1:
            self.y = float(value)
        else:
            raise IndexError("Index must be in (0,1)")

    def __repr__(self):
        return "Vec2(%g,%g)" % (self.x, self.y)

    def __copy__(self):
        return Vec2(self.x, self.y)

    copy = __copy__

    def __neg__(self):
        return Vec2(-self.x, -self.y)

    def __abs__(self):
        return Vec2(abs(self.x), abs(self.y))

    def __add__(self, o):
        ox, oy = o
        return Vec2(self.x + ox, self.y + oy)

    __radd__ = __add__

    def __sub__(self, o):
        return Vec2(self.x - o[0], self.y - o[1])

    def __rsub__(self, o):
        return Vec2(o[0] - self.x, o[1] - self.y)

    def __mul__(self, s):
        if isinstance(s, Number):
            return Vec2(s * self.x, s * self.y)
        return self.dot(s)

    __rmul__ = __mul__

    def __truediv__(self, s):
        if isinstance(s, Number):
            return Vec2(self.x / s, self.y / s)
        raise ValueError("Ambiguous operation")

    def __eq__(self, o):
        try:
            return self.x == o[0] and self.y == o[1]
        except Exception:
            return False

    def __ne__(self, o):
        return not self.__eq__(o)

    __hash__ = None

    def dot(self, o):
        ox, oy = o
        return self.x * float(ox) + self.y * float(oy)

    def cross(self, o):
        # vector x scalar -> vector ; vector x vector -> scalar (common.py:204-212)
        if isinstance(o, Number):
            return Vec2(float(o) * self.y, -float(o) * self.x)
        ox, oy = o
        return self.x * float(oy) - self.y * float(ox)

    @property
    def length_squared(self):
        return self.x ** 2 + self.y ** 2

    @property
    def length(self):
        return math.sqrt(self.x ** 2 + self.y ** 2)

    @property
    def normalized(self):
        n = self.length
        if n < EPSILON:
            return Vec2(self.x, self.y)
        inv = 1.0 / n
        return Vec2(self.x * inv, self.y * inv)

    @property
    def valid(self):
        return not (math.isnan(self.x) or math.isnan(self.y))


class Transform(tuple):
    """A rigid transform as `Body.transform` hands it out (common.py:618-706: `.position`, `.angle`, `xf * v`, `xf.mul_t(v)`),
    stored the way this package keeps transforms everywhere: the tuple (px, py, cos, sin)."""
    __slots__ = ()

    def __new__(cls, position=(0, 0), angle=0.0, cos_sin=None):
        c, s = cos_sin if cos_sin is not None else (math.cos(angle), math.sin(angle))
        return tuple.__new__(cls, (float(position[0]), float(position[1]), float(c), float(s)))

    def __repr__(self):
        return "Transform(position=%s, angle=%g)" % (self.position, self.angle)

    @property
    def position(self):
        return Vec2(self[0], self[1])

    @property
    def angle(self):
        return math.atan2(self[3], self[2])

    def __mul__(self, v):
        x, y = v
        return Vec2(self[0] + self[2] * x + (-self[3]) * y, self[1] + self[3] * x + self[2] * y)

    def mul_t(self, v):
        x, y = v[0] - self[0], v[1] - self[1]
        return Vec2(self[2] * x + self[3] * y, (-self[3]) * x + self[2] * y)


class AABB(object):
    """An axis-aligned box as callers of `World.query_aabb` build it (common.py:707-870: `AABB(lower_bound, upper_bound)`): a value
    object over four floats with the reference's read-outs."""
    __slots__ = ("lx", "ly", "ux", "uy")

    def __init__(self, lower_bound=(0, 0), upper_bound=(0, 0)):
        (self.lx, self.ly), (self.ux, self.uy) = (float(v) for v in lower_bound), (float(v) for v in upper_bound)

    def __repr__(self):
        return "AABB(lower_bound=%s, upper_bound=%s)" % (self.lower_bound, self.upper_bound)

    def __copy__(self):
        return AABB((self.lx, self.ly), (self.ux, self.uy))

    @property
    def lower_bound(self):
        return Vec2(self.lx, self.ly)

    @lower_bound.setter
    def lower_bound(self, v):
        self.lx, self.ly = (float(x) for x in v)

    @property
    def upper_bound(self):
        return Vec2(self.ux, self.uy)

    @upper_bound.setter
    def upper_bound(self, v):
        self.ux, self.uy = (float(x) for x in v)

    @property
    def valid(self):
        return self.ux - self.lx >= 0.0 and self.uy - self.ly >= 0.0 and all(is_valid_float(v) for v in (self.lx, self.ly, self.ux, self.uy))

    @property
    def center(self):
        return Vec2(0.5 * (self.lx + self.ux), 0.5 * (self.ly + self.uy))

    @property
    def extents(self):
        return Vec2(0.5 * (self.ux - self.lx), 0.5 * (self.uy - self.ly))

    @property
    def perimeter(self):
        return 2.0 * ((self.ux - self.lx) + (self.uy - self.ly))

    def __add__(self, other):
        return AABB((min(self.lx, other.lx), min(self.ly, other.ly)), (max(self.ux, other.ux), max(self.uy, other.uy)))

    def contains(self, other):
        return self.lx <= other.lx and self.ly <= other.ly and other.ux <= self.ux and other.uy <= self.uy

    def test_overlap(self, other):
        return not (other.lx - self.ux > 0.0 or other.ly - self.uy > 0.0 or self.lx - other.ux > 0.0 or self.ly - other.uy > 0.0)


def scalar_cross(s, v):
    """scalar x vector (common.py:195-202)."""
    return Vec2(-float(s) * v[1], float(s) * v[0])


def clamp(value, low, high):
    return max(low, min(value, high))


def is_valid_float(x):
    return not math.isnan(x)


def rot(angle):
    """(cos, sin) of a body angle, libm double as the reference (common.py:495-503)."""
    return math.cos(angle), math.sin(angle)


def xf_apply(xf, v):
    """xf = (px, py, c, s); p + R v with the reference's operation order (common.py:677-685)."""
    px, py, c, s = xf
    return (px + c * v[0] + (-s) * v[1], py + s * v[0] + c * v[1])
