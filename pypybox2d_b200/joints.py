"""Joints of the island solver: the drop-in surface of `pypybox2d/joints/`.

`RevoluteJoint` (`joints/revolute.py`), `DistanceJoint` (`joints/distance.py`), `WeldJoint` (`joints/weld.py`) and
`PrismaticJoint` (`joints/prismatic.py`), `RopeJoint` (`joints/rope.py`), `FrictionJoint` (`joints/friction.py`), `WheelJoint`
(`joints/wheel.py`), `PulleyJoint` (`joints/pulley.py`) and `MouseJoint` (`joints/mouse.py`) with the
reference's constructor keywords, read-backs and setters.  A joint object only holds its definition: one row of the device
arrays `j_def` (type, bodies, flags), `j_par` (parameters, layout below) and `j_imp` (accumulated impulses, mirrored back
after every step).  The constraint itself -- `_init_velocity_constraints`, `_solve_velocity_constraints`,
`_solve_position_constraints` -- runs inside the solve kernels (`csrc/b2b_joints.cuh`), there is no host implementation.

`j_par` rows (reals):
  every type  [0:2] local_anchor_a, [2:4] local_anchor_b
  revolute    [4] reference_angle, [5] lower_angle, [6] upper_angle, [7] max_motor_torque, [8] motor_speed
  distance    [4] length, [5] frequency, [6] damping_ratio
  weld        [4] reference_angle
  prismatic   [4] reference_angle, [5] lower_limit, [6] upper_limit, [7] max_motor_force, [8] motor_speed, [9:11] local_x_axis
  rope        [4] max_length
  friction    [4] max_force, [5] max_torque
  wheel       [4:6] local_x_axis, [6] frequency, [7] damping_ratio, [8] max_motor_torque, [9] motor_speed
  pulley      [4:6] ground_anchor_a, [6:8] ground_anchor_b, [8] ratio, [9] constant = length_a + ratio * length_b
  mouse       (one body: j_def body_a = -1) [4:6] target, [6] max_force, [7] frequency, [8] damping_ratio, [9] Body.mass
GearJoint (a constraint between two other joints) raises NotImplementedError.
"""
import math

from . import schema
from .common import Vec2
from .settings import LINEAR_SLOP

INACTIVE_LIMIT, AT_LOWER_LIMIT, AT_UPPER_LIMIT, EQUAL_LIMITS = range(4)   # joints/joint.py:36


def _world_point(body, lp):
    # Transform.__mul__ (common.py:677-685): position.x + col1.x * v.x + col2.x * v.y, left to right
    x, y, c, s = body._xf
    return Vec2(x + c * lp.x + (-s) * lp.y, y + s * lp.x + c * lp.y)


class Joint(object):
    """joints/joint.py:39-112"""
    _type = 0

    def __init__(self, body_a, body_b, collide_connected):
        if body_a and body_b:
            if body_a._world != body_b.world:
                raise ValueError('Both bodies must be in the same world.')
            if body_a._world is None or body_b._world is None:
                raise ValueError('Both bodies must be in the same world.')
        self._body_a = body_a
        self._body_b = body_b
        self._collide_connected = bool(collide_connected)
        self._impulse3 = [0.0, 0.0, 0.0]
        self._motor_impulse = 0.0
        self._limit_state = INACTIVE_LIMIT
        self._limit_enabled = False
        self._motor_enabled = False
        self._world = None

    body_a = property(lambda self: self._body_a)
    body_b = property(lambda self: self._body_b)
    bodies = property(lambda self: [self._body_a, self._body_b])
    collide_connected = property(lambda self: self._collide_connected)

    def other_body(self, body):
        return self._body_b if body == self._body_a else self._body_a

    @property
    def active(self):
        return self._body_a.active and self._body_b.active

    @property
    def local_anchor_a(self):
        return Vec2(*self._local_anchor_a)

    @property
    def local_anchor_b(self):
        return Vec2(*self._local_anchor_b)

    @property
    def anchor_a(self):
        return self._body_a.get_world_point(self._local_anchor_a)

    @property
    def anchor_b(self):
        return self._body_b.get_world_point(self._local_anchor_b)

    def get_reaction_force(self, inv_dt):
        return Vec2(inv_dt * self._impulse3[0], inv_dt * self._impulse3[1])

    def get_reaction_torque(self, inv_dt):
        return inv_dt * self._impulse3[2]

    # -- device rows -------------------------------------------------------------------------------------------------
    def _flags(self):
        f = 0
        if self._collide_connected:
            f |= schema.JOINT_COLLIDE_CONNECTED
        if self._limit_enabled:
            f |= schema.JOINT_LIMIT_ENABLED
        if self._motor_enabled:
            f |= schema.JOINT_MOTOR_ENABLED
        return f | (self._limit_state << schema.JOINT_LIMIT_SHIFT)

    def _par(self):
        raise NotImplementedError

    def _imp(self):
        return (self._impulse3[0], self._impulse3[1], self._impulse3[2], self._motor_impulse)

    def _read_back(self, flags, imp):
        self._limit_state = (int(flags) >> schema.JOINT_LIMIT_SHIFT) & 3
        self._impulse3 = [float(imp[0]), float(imp[1]), float(imp[2])]
        self._motor_impulse = float(imp[3])

    def _changed(self, wake=True):
        if wake:
            for body in self.bodies:
                if body is not None:
                    body.awake = True
        if self._world is not None:
            self._world._joints_dirty = True
            self._world._host_dirty = True


class RevoluteJoint(Joint):
    """joints/revolute.py:33-453"""
    _type = schema.JOINT_REVOLUTE

    def __init__(self, body_a, body_b, anchor=(0, 0), reference_angle=None, lower_angle=0.0, upper_angle=0.0,
                 max_motor_torque=0.0, motor_speed=0.0, limit_enabled=False, motor_enabled=False, local_anchor_a=None,
                 local_anchor_b=None, collide_connected=False):
        if body_a is None or body_b is None:
            raise ValueError('body_a and body_b must be set')
        Joint.__init__(self, body_a, body_b, collide_connected)
        self._local_anchor_a = Vec2(*local_anchor_a) if local_anchor_a is not None else body_a.get_local_point(anchor)
        self._local_anchor_b = Vec2(*local_anchor_b) if local_anchor_b is not None else body_b.get_local_point(anchor)
        self._reference_angle = float(reference_angle) if reference_angle is not None else body_b.angle - body_a.angle
        self._lower_angle = float(lower_angle)
        self._upper_angle = float(upper_angle)
        self._max_motor_torque = float(max_motor_torque)
        self._limit_enabled = bool(limit_enabled)
        self._motor_enabled = bool(motor_enabled)
        self._motor_speed = float(motor_speed)

    def __copy__(self):
        return RevoluteJoint(self.body_a, self.body_b, None, self._reference_angle, self._lower_angle, self._upper_angle,
                             self._max_motor_torque, self._motor_speed, self._limit_enabled, self._motor_enabled,
                             self._local_anchor_a, self._local_anchor_b, self._collide_connected)

    def _par(self):
        return [self._local_anchor_a.x, self._local_anchor_a.y, self._local_anchor_b.x, self._local_anchor_b.y,
                self._reference_angle, self._lower_angle, self._upper_angle, self._max_motor_torque, self._motor_speed]

    @property
    def joint_angle(self):
        return self._body_b._a - self._body_a._a - self._reference_angle

    @property
    def joint_speed(self):
        return self._body_b._angular_velocity - self._body_a._angular_velocity

    @property
    def motor_speed(self):
        return self._motor_speed

    @motor_speed.setter
    def motor_speed(self, speed):
        self._motor_speed = float(speed)
        self._changed()

    @property
    def limit_enabled(self):
        return self._limit_enabled

    @limit_enabled.setter
    def limit_enabled(self, limit_enabled):
        if limit_enabled != self._limit_enabled:
            self._limit_enabled = bool(limit_enabled)
            self._impulse3[2] = 0.0
            self._changed()

    @property
    def motor_enabled(self):
        return self._motor_enabled

    @motor_enabled.setter
    def motor_enabled(self, motor_enabled):
        if motor_enabled != self._motor_enabled:
            self._motor_enabled = bool(motor_enabled)
            self._changed()

    @property
    def lower_angle(self):
        return self._lower_angle

    @lower_angle.setter
    def lower_angle(self, lower_angle):
        if self._lower_angle != lower_angle:
            self._impulse3[2] = 0.0
            self._lower_angle = float(lower_angle)
            self._changed()

    @property
    def upper_angle(self):
        return self._upper_angle

    @upper_angle.setter
    def upper_angle(self, upper_angle):
        if self._upper_angle != upper_angle:
            self._impulse3[2] = 0.0
            self._upper_angle = float(upper_angle)
            self._changed()

    @property
    def max_motor_torque(self):
        return self._max_motor_torque

    @max_motor_torque.setter
    def max_motor_torque(self, torque):
        self._max_motor_torque = float(torque)
        self._changed()

    def get_motor_torque(self, inv_dt):
        return inv_dt * self._motor_impulse


class DistanceJoint(Joint):
    """joints/distance.py:33-311"""
    _type = schema.JOINT_DISTANCE

    def __init__(self, body_a, body_b, anchor_a=(0, 0), anchor_b=(0, 0), frequency=0.0, damping_ratio=0.0, length=None,
                 collide_connected=False, local_anchor_a=None, local_anchor_b=None):
        if body_a is None or body_b is None:
            raise ValueError('body_a and body_b must be set')
        Joint.__init__(self, body_a, body_b, collide_connected)
        self._local_anchor_a = Vec2(*local_anchor_a) if local_anchor_a is not None else body_a.get_local_point(anchor_a)
        self._local_anchor_b = Vec2(*local_anchor_b) if local_anchor_b is not None else body_b.get_local_point(anchor_b)
        if length is not None:
            self._length = float(length)
        else:
            pa, pb = _world_point(body_a, self._local_anchor_a), _world_point(body_b, self._local_anchor_b)
            dx, dy = pb.x - pa.x, pb.y - pa.y
            self._length = math.sqrt(dx ** 2 + dy ** 2)
        if self._length <= LINEAR_SLOP:
            raise ValueError('length must be > 0 and not a very small number')
        self._frequency = float(frequency)
        self._damping_ratio = float(damping_ratio)

    def __copy__(self):
        return DistanceJoint(self.body_a, self.body_b, self.anchor_a, self.anchor_b, self._frequency, self._damping_ratio,
                             self._length, self._collide_connected, self._local_anchor_a, self._local_anchor_b)

    def _par(self):
        return [self._local_anchor_a.x, self._local_anchor_a.y, self._local_anchor_b.x, self._local_anchor_b.y,
                self._length, self._frequency, self._damping_ratio]

    @property
    def length(self):
        return self._length

    @length.setter
    def length(self, length):
        if length <= LINEAR_SLOP:
            raise ValueError('Length must be > 0 and not a very small number')
        self._length = float(length)
        self._changed(wake=False)

    @property
    def frequency(self):
        return self._frequency

    @frequency.setter
    def frequency(self, frequency):
        self._frequency = float(frequency)
        self._changed(wake=False)

    @property
    def damping_ratio(self):
        return self._damping_ratio

    @damping_ratio.setter
    def damping_ratio(self, damping_ratio):
        self._damping_ratio = float(damping_ratio)
        self._changed(wake=False)

    def get_reaction_force(self, inv_dt):
        # distance.py: inv_dt * impulse * u; the unit vector is a solver temporary of the last step (device `j_tmp`)
        u = getattr(self, "_u", (0.0, 0.0))
        return Vec2(inv_dt * self._impulse3[0] * u[0], inv_dt * self._impulse3[0] * u[1])

    def get_reaction_torque(self, inv_dt):
        return 0.0


class WeldJoint(Joint):
    """joints/weld.py:33-257"""
    _type = schema.JOINT_WELD

    def __init__(self, body_a, body_b, anchor=None, reference_angle=None, local_anchor_a=None, local_anchor_b=None,
                 collide_connected=False):
        if body_a is None or body_b is None:
            raise ValueError('body_a and body_b must be set')
        Joint.__init__(self, body_a, body_b, collide_connected)
        if anchor is not None:
            self._local_anchor_a = body_a.get_local_point(anchor)
            self._local_anchor_b = body_b.get_local_point(anchor)
        else:
            self._local_anchor_a = Vec2(*local_anchor_a)
            self._local_anchor_b = Vec2(*local_anchor_b)
        self._reference_angle = float(reference_angle) if reference_angle is not None else body_b.angle - body_a.angle

    def __copy__(self):
        return WeldJoint(self._body_a, self._body_b, None, self._reference_angle, self._local_anchor_a, self._local_anchor_b,
                         self._collide_connected)

    def _par(self):
        return [self._local_anchor_a.x, self._local_anchor_a.y, self._local_anchor_b.x, self._local_anchor_b.y,
                self._reference_angle]


class PrismaticJoint(Joint):
    """joints/prismatic.py:33-513"""
    _type = schema.JOINT_PRISMATIC

    def __init__(self, body_a, body_b, anchor=(0, 0), axis=(0, 1), lower_limit=0.0, upper_limit=0.0, max_motor_force=0.0,
                 max_motor_torque=0.0, motor_speed=0.0, reference_angle=None, limit_enabled=False, motor_enabled=False,
                 collide_connected=False, local_anchor_a=None, local_anchor_b=None, local_x_axis=None):
        if body_a is None or body_b is None:
            raise ValueError('body_a and body_b must be set')
        Joint.__init__(self, body_a, body_b, collide_connected)
        self._local_anchor_a = Vec2(*local_anchor_a) if local_anchor_a is not None else body_a.get_local_point(anchor)
        self._local_anchor_b = Vec2(*local_anchor_b) if local_anchor_b is not None else body_b.get_local_point(anchor)
        self._reference_angle = float(reference_angle) if reference_angle is not None else body_b.angle - body_a.angle
        self._local_x_axis = Vec2(*local_x_axis) if local_x_axis is not None else body_a.get_local_vector(axis)
        self._lower_limit = float(lower_limit)
        self._upper_limit = float(upper_limit)
        self._max_motor_force = float(max_motor_force)
        self._limit_enabled = bool(limit_enabled)
        self._motor_enabled = bool(motor_enabled)
        self._motor_speed = float(motor_speed)

    def _par(self):
        return [self._local_anchor_a.x, self._local_anchor_a.y, self._local_anchor_b.x, self._local_anchor_b.y,
                self._reference_angle, self._lower_limit, self._upper_limit, self._max_motor_force, self._motor_speed,
                self._local_x_axis.x, self._local_x_axis.y]

    @property
    def axis(self):
        return self._body_a.get_world_vector(self._local_x_axis)

    @property
    def joint_translation(self):
        d = self.anchor_b - self.anchor_a
        return d.dot(self.axis)

    @property
    def limit_enabled(self):
        return self._limit_enabled

    @limit_enabled.setter
    def limit_enabled(self, limit_enabled):
        if limit_enabled != self._limit_enabled:
            self._limit_enabled = bool(limit_enabled)
            self._impulse3[2] = 0.0
            self._changed()

    @property
    def lower_limit(self):
        return self._lower_limit

    @lower_limit.setter
    def lower_limit(self, lower_limit):
        if self._lower_limit != lower_limit:
            self._impulse3[2] = 0.0
            self._lower_limit = float(lower_limit)
            self._changed()

    @property
    def upper_limit(self):
        return self._upper_limit

    @upper_limit.setter
    def upper_limit(self, upper_limit):
        if self._upper_limit != upper_limit:
            self._impulse3[2] = 0.0
            self._upper_limit = float(upper_limit)
            self._changed()

    @property
    def motor_enabled(self):
        return self._motor_enabled

    @motor_enabled.setter
    def motor_enabled(self, motor_enabled):
        if motor_enabled != self._motor_enabled:
            self._motor_enabled = bool(motor_enabled)
            self._changed()

    @property
    def motor_speed(self):
        return self._motor_speed

    @motor_speed.setter
    def motor_speed(self, speed):
        self._motor_speed = float(speed)
        self._changed()

    @property
    def max_motor_force(self):
        return self._max_motor_force

    @max_motor_force.setter
    def max_motor_force(self, force):
        self._max_motor_force = float(force)
        self._changed()

    def get_motor_force(self, inv_dt):
        return inv_dt * self._motor_impulse

    def get_reaction_torque(self, inv_dt):
        return inv_dt * self._impulse3[1]

    def get_reaction_force(self, inv_dt):
        # prismatic.py:108-110: inv_dt * (impulse.x * perp + (motor_impulse + impulse.z) * axis); perp and axis are solver
        # temporaries of the last step (device `j_tmp`)
        ax, px = getattr(self, "_axis_perp", ((0.0, 0.0), (0.0, 0.0)))
        k = self._motor_impulse + self._impulse3[2]
        return Vec2(inv_dt * (self._impulse3[0] * px[0] + k * ax[0]), inv_dt * (self._impulse3[0] * px[1] + k * ax[1]))


class RopeJoint(Joint):
    """joints/rope.py:33-248: a maximum distance between two anchor points."""
    _type = schema.JOINT_ROPE

    def __init__(self, body_a, body_b, local_anchor_a=(0, 0), local_anchor_b=(0, 0), max_length=0.0, collide_connected=False):
        if body_a is None or body_b is None:
            raise ValueError('body_a and body_b must be set')
        Joint.__init__(self, body_a, body_b, collide_connected)
        self._local_anchor_a = Vec2(*local_anchor_a)
        self._local_anchor_b = Vec2(*local_anchor_b)
        self._max_length = float(max_length)

    def __copy__(self):
        return RopeJoint(self._body_a, self._body_b, self._local_anchor_a, self._local_anchor_b, self._max_length,
                         self._collide_connected)

    def _par(self):
        return [self._local_anchor_a.x, self._local_anchor_a.y, self._local_anchor_b.x, self._local_anchor_b.y, self._max_length]

    max_length = property(lambda self: self._max_length)
    limit_state = property(lambda self: self._limit_state)

    def get_reaction_force(self, inv_dt):
        u = getattr(self, "_u", (0.0, 0.0))
        return Vec2((inv_dt * self._impulse3[0]) * u[0], (inv_dt * self._impulse3[0]) * u[1])

    def get_reaction_torque(self, inv_dt):
        return 0.0


class FrictionJoint(Joint):
    """joints/friction.py:33-233 (as the port behaves: its velocity solve stores its result into `positions`, so only the warm
    start of the accumulated impulses acts on the bodies)."""
    _type = schema.JOINT_FRICTION

    def __init__(self, body_a, body_b, local_anchor_a=(0, 0), local_anchor_b=(0, 0), max_force=0.0, max_torque=0.0,
                 collide_connected=False):
        if body_a is None or body_b is None:
            raise ValueError('body_a and body_b must be set')
        Joint.__init__(self, body_a, body_b, collide_connected)
        self._local_anchor_a = Vec2(*local_anchor_a)
        self._local_anchor_b = Vec2(*local_anchor_b)
        self._max_force = float(max_force)
        self._max_torque = float(max_torque)

    def __copy__(self):
        return FrictionJoint(self._body_a, self._body_b, self._local_anchor_a, self._local_anchor_b, self._max_force,
                             self._max_torque, self._collide_connected)

    def _par(self):
        return [self._local_anchor_a.x, self._local_anchor_a.y, self._local_anchor_b.x, self._local_anchor_b.y,
                self._max_force, self._max_torque]

    @property
    def max_force(self):
        return self._max_force

    @max_force.setter
    def max_force(self, force):
        self._max_force = float(force)
        self._changed(wake=False)

    @property
    def max_torque(self):
        return self._max_torque

    @max_torque.setter
    def max_torque(self, torque):
        self._max_torque = float(torque)
        self._changed(wake=False)


class WheelJoint(Joint):
    """joints/wheel.py:33-417: a point of body_b on a line fixed in body_a, with a spring along it and a rotational motor."""
    _type = schema.JOINT_WHEEL

    def __init__(self, body_a, body_b, anchor=None, axis=None, frequency=2.0, damping_ratio=0.7, motor_enabled=False,
                 motor_speed=0.0, max_motor_torque=0.0, local_anchor_a=None, local_anchor_b=None, local_x_axis=None,
                 collide_connected=False):
        if body_a is None or body_b is None:
            raise ValueError('body_a and body_b must be set')
        Joint.__init__(self, body_a, body_b, collide_connected)
        self._local_anchor_a = Vec2(*local_anchor_a) if local_anchor_a is not None else body_a.get_local_point(anchor)
        self._local_anchor_b = Vec2(*local_anchor_b) if local_anchor_b is not None else body_b.get_local_point(anchor)
        self._local_x_axis = Vec2(*local_x_axis) if local_x_axis is not None else body_a.get_local_vector(axis)
        self._frequency = float(frequency)
        self._damping_ratio = float(damping_ratio)
        self._max_motor_torque = float(max_motor_torque)
        self._motor_speed = float(motor_speed)
        self._motor_enabled = bool(motor_enabled)

    def __copy__(self):
        return WheelJoint(self._body_a, self._body_b, None, None, self._frequency, self._damping_ratio, self._motor_enabled,
                          self._motor_speed, self._max_motor_torque, self._local_anchor_a, self._local_anchor_b,
                          self._local_x_axis, self._collide_connected)

    def _par(self):
        return [self._local_anchor_a.x, self._local_anchor_a.y, self._local_anchor_b.x, self._local_anchor_b.y,
                self._local_x_axis.x, self._local_x_axis.y, self._frequency, self._damping_ratio, self._max_motor_torque,
                self._motor_speed]

    # accumulated impulses: _impulse (point to line), _spring_impulse, _motor_impulse ride in j_imp[0:3]
    def _imp(self):
        return (self._impulse3[0], self._impulse3[1], self._impulse3[2], 0.0)

    def _read_back(self, flags, imp):
        Joint._read_back(self, flags, imp)
        self._motor_impulse = float(imp[2])

    @property
    def joint_speed(self):
        return self._body_b._angular_velocity - self._body_a._angular_velocity

    @property
    def motor_enabled(self):
        return self._motor_enabled

    @motor_enabled.setter
    def motor_enabled(self, motor_enabled):
        self._motor_enabled = bool(motor_enabled)
        self._changed()

    @property
    def motor_speed(self):
        return self._motor_speed

    @motor_speed.setter
    def motor_speed(self, speed):
        self._motor_speed = float(speed)
        self._changed()

    @property
    def max_motor_torque(self):
        return self._max_motor_torque

    @max_motor_torque.setter
    def max_motor_torque(self, torque):
        self._max_motor_torque = float(torque)
        self._changed()

    @property
    def spring_frequency(self):
        return self._frequency

    @spring_frequency.setter
    def spring_frequency(self, frequency):
        self._frequency = float(frequency)
        self._changed(wake=False)

    @property
    def spring_damping_ratio(self):
        return self._damping_ratio

    @spring_damping_ratio.setter
    def spring_damping_ratio(self, damping_ratio):
        self._damping_ratio = float(damping_ratio)
        self._changed(wake=False)

    def get_motor_torque(self, inv_dt):
        return inv_dt * self._impulse3[2]

    def get_reaction_torque(self, inv_dt):
        return inv_dt * self._impulse3[2]


class PulleyJoint(Joint):
    """joints/pulley.py:33-320: length_a + ratio * length_b stays constant; the ground anchors are world points."""
    _type = schema.JOINT_PULLEY

    def __init__(self, body_a, body_b, ground_anchor_a=(0, 0), ground_anchor_b=(0, 0), anchor_a=(0, 0), anchor_b=(0, 0), ratio=1.0,
                 length_a=None, length_b=None, collide_connected=False, local_anchor_a=None, local_anchor_b=None):
        from .settings import EPSILON
        if body_a is None or body_b is None:
            raise ValueError('body_a and body_b must be set')
        Joint.__init__(self, body_a, body_b, collide_connected)
        if ratio == 0.0 or ratio < EPSILON:
            raise ValueError('Ratio must be positive, non-zero and not too small')
        self._local_anchor_a = body_a.get_local_point(anchor_a) if local_anchor_a is None else Vec2(*local_anchor_a)
        self._local_anchor_b = body_b.get_local_point(anchor_b) if local_anchor_b is None else Vec2(*local_anchor_b)
        self._ground_anchor_a = Vec2(*ground_anchor_a)
        self._ground_anchor_b = Vec2(*ground_anchor_b)
        if length_a is None:
            dx, dy = anchor_a[0] - self._ground_anchor_a.x, anchor_a[1] - self._ground_anchor_a.y
            length_a = math.sqrt(dx ** 2 + dy ** 2)
        if length_b is None:
            dx, dy = anchor_b[0] - self._ground_anchor_b.x, anchor_b[1] - self._ground_anchor_b.y
            length_b = math.sqrt(dx ** 2 + dy ** 2)
        self._length_a = length_a
        self._length_b = length_b
        self._constant = length_a + ratio * length_b
        self._ratio = ratio

    def __copy__(self):
        return PulleyJoint(self._body_a, self._body_b, self._ground_anchor_a, self._ground_anchor_b, None, None, self._ratio,
                           self._length_a, self._length_b, self._collide_connected, self._local_anchor_a, self._local_anchor_b)

    def _par(self):
        return [self._local_anchor_a.x, self._local_anchor_a.y, self._local_anchor_b.x, self._local_anchor_b.y,
                self._ground_anchor_a.x, self._ground_anchor_a.y, self._ground_anchor_b.x, self._ground_anchor_b.y,
                self._ratio, self._constant]

    ground_anchor_a = property(lambda self: Vec2(*self._ground_anchor_a))
    ground_anchor_b = property(lambda self: Vec2(*self._ground_anchor_b))
    ratio = property(lambda self: self._ratio)

    @property
    def length_a(self):
        p = self.anchor_a
        return math.hypot(p.x - self._ground_anchor_a.x, p.y - self._ground_anchor_a.y)

    @property
    def length_b(self):
        p = self.anchor_b
        return math.hypot(p.x - self._ground_anchor_b.x, p.y - self._ground_anchor_b.y)

    def get_reaction_force(self, inv_dt):
        u = getattr(self, "_ub", (0.0, 0.0))
        return Vec2(inv_dt * (self._impulse3[0] * u[0]), inv_dt * (self._impulse3[0] * u[1]))

    def get_reaction_torque(self, inv_dt):
        return 0.0


class MouseJoint(Joint):
    """joints/mouse.py:33-229: a soft constraint that pulls a point of ONE body to a world target (body_a is None)."""
    _type = schema.JOINT_MOUSE

    def __init__(self, body, target=(0, 0), max_force=0.0, frequency=5.0, damping_ratio=0.7):
        from .common import is_valid_float
        if body is None:
            raise ValueError('body must be set')
        target = Vec2(*target)
        if not (is_valid_float(target.x) and is_valid_float(target.y)):
            raise ValueError('Invalid target')
        if not is_valid_float(max_force) or max_force < 0.0:
            raise ValueError('Invalid maximum force')
        if not is_valid_float(frequency) or frequency < 0.0:
            raise ValueError('Invalid frequency')
        if not is_valid_float(damping_ratio) or damping_ratio < 0.0:
            raise ValueError('Invalid damping ratio')
        Joint.__init__(self, None, body, False)
        self._target = target
        self._local_anchor_a = Vec2()
        self._local_anchor_b = body.get_local_point(target)
        self._max_force = float(max_force)
        self._frequency = float(frequency)
        self._damping_ratio = float(damping_ratio)

    def __copy__(self):
        return MouseJoint(self._body_b, self._target, self._max_force, self._frequency, self._damping_ratio)

    def _par(self):
        return [0.0, 0.0, self._local_anchor_b.x, self._local_anchor_b.y, self._target.x, self._target.y, self._max_force,
                self._frequency, self._damping_ratio, self._body_b.mass]

    active = property(lambda self: self._body_b.active)
    anchor_a = property(lambda self: Vec2(*self._target))

    @property
    def target(self):
        return Vec2(*self._target)

    @target.setter
    def target(self, target):
        if not self._body_b.awake:
            self._body_b.awake = True
        self._target = Vec2(*target)
        self._changed(wake=False)

    @property
    def max_force(self):
        return self._max_force

    @max_force.setter
    def max_force(self, max_force):
        self._max_force = float(max_force)
        self._changed(wake=False)

    @property
    def frequency(self):
        return self._frequency

    @frequency.setter
    def frequency(self, frequency):
        self._frequency = float(frequency)
        self._changed(wake=False)

    @property
    def damping_ratio(self):
        return self._damping_ratio

    @damping_ratio.setter
    def damping_ratio(self, damping_ratio):
        self._damping_ratio = float(damping_ratio)
        self._changed(wake=False)

    def get_reaction_torque(self, inv_dt):
        return 0.0


def _unsupported(name):
    def ctor(*args, **kwargs):
        raise NotImplementedError("%s is not supported on the batched step path (built: revolute, distance, weld, prismatic, rope, friction, wheel, pulley, mouse)" % name)
    ctor.__name__ = name
    return ctor


GearJoint = _unsupported("GearJoint")


def pack_joints(world, arrays, w, body_index):
    """World.joints -> row `w` of the joint arrays; Body.joints -> the CSR rows b_jadj / j_adj (append order of
    World.add_joint, world.py:268-273: body_a's list first, then body_b's)."""
    A = arrays
    joints = world.joints
    nj = len(joints)
    if nj > (A["j_def"].shape[1] if nj else 0) and nj > 0:
        raise ValueError("world exceeds the batch's joint capacity (Caps.nj)")
    A["w_joint"][w] = (nj, 0, 0, 0)
    jidx = {id(j): k for k, j in enumerate(joints)}
    for k, j in enumerate(joints):
        A["j_def"][w, k] = (j._type, body_index[id(j._body_a)] if j._body_a is not None else -1, body_index[id(j._body_b)], j._flags())
        par = j._par()
        A["j_par"][w, k, :len(par)] = par
        A["j_imp"][w, k] = j._imp()
    pos = 0
    for i, b in enumerate(world.bodies):
        row = [jidx[id(j)] for j in b.joints]
        A["b_jadj"][w, i] = (pos, len(row))
        flat = A["j_adj"][w].reshape(-1)
        flat[pos:pos + len(row)] = row
        pos += len(row)
