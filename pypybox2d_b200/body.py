"""Rigid body facade (host side).

Same constructor keywords and read-backs as the reference's `Body`
(`pypybox2d/body.py:39-44`, properties `:415-561,792-812`).  The object is a host mirror:
`World.step` packs all bodies into device arrays, runs the CUDA step and writes the new
state back into these attributes.  Mass aggregation follows `body.py:606-660` in the same
operation order so inverse masses and inertias are bit-identical to the reference's.
"""
import math
from copy import copy

from . import shapes as _shapes
from .common import Vec2, Transform, LockedError, is_valid_float
from .fixture import Fixture


class Body(object):
    STATIC = 0
    DYNAMIC = 1
    KINEMATIC = 2

    def __init__(self, world=None, position=(0, 0), angle=0.0, linear_velocity=(0, 0),
                 angular_velocity=0.0, linear_damping=0.0, angular_damping=0.0,
                 allow_sleep=True, awake=True, fixed_rotation=False, bullet=False,
                 type=STATIC, active=True, gravity_scale=1.0, check_args=True,
                 shapes=[], fixtures=[], shape_fixture=None, user_data=None,
                 copy_fixtures=True):
        position = Vec2(*position)
        linear_velocity = Vec2(*linear_velocity)
        if check_args:
            if not position.valid:
                raise ValueError("Invalid position")
            if not linear_velocity.valid:
                raise ValueError("Invalid linear velocity")
            if not is_valid_float(angle):
                raise ValueError("Invalid angle")
            if not is_valid_float(angular_velocity):
                raise ValueError("Invalid angular velocity")
            if type not in (Body.STATIC, Body.DYNAMIC, Body.KINEMATIC):
                raise ValueError("Invalid body type")

        angle = float(angle)
        self._xf = (position.x, position.y, math.cos(angle), math.sin(angle))
        self._local_center = Vec2()
        self._a = self._a0 = angle
        self._c0 = Vec2(*self._xf_mul(self._local_center))
        self._c = copy(self._c0)
        self._force = Vec2()
        self._torque = 0.0
        self._I = 0.0
        self._invI = 0.0
        self._sleep_time = 0.0
        self.fixtures = []
        self.joints = []
        self._controllers = []
        self.user_data = user_data
        self._world = world
        self._linear_velocity = linear_velocity
        self._angular_velocity = float(angular_velocity)
        self._linear_damping = float(linear_damping)
        self._angular_damping = float(angular_damping)
        self._gravity_scale = float(gravity_scale)
        self._island_index = None
        self._awake = bool(awake)
        self._allow_sleep = bool(allow_sleep)
        self._fixed_rotation = bool(fixed_rotation)
        self._bullet = bool(bullet)
        self._type = type
        self._active = bool(active)      # an inactive body has no proxies: it collides with nothing, is in no island, never moves
        if self._type == Body.DYNAMIC:
            self._mass = 1.0
            self._inv_mass = 1.0
        else:
            self._mass = 0.0
            self._inv_mass = 0.0

        if shapes is not None:
            self.create_fixtures(shapes, shape_fixture)
        if fixtures is not None:
            if not isinstance(fixtures, (list, tuple)):
                fixtures = [fixtures]
            for fixture in fixtures:
                self.attach_fixture(copy(fixture) if copy_fixtures else fixture)

    def __copy__(self):
        """A detached copy with copies of the fixtures (body.py:118-140); World.create_body(body=...) uses it."""
        px, py, c, s = self._xf
        new_body = Body(position=(px, py), angle=self._a, linear_velocity=self._linear_velocity,
                        angular_velocity=self._angular_velocity, linear_damping=self._linear_damping,
                        angular_damping=self._angular_damping, allow_sleep=self._allow_sleep, awake=self._awake,
                        fixed_rotation=self._fixed_rotation, bullet=self._bullet, type=self._type, active=self._active,
                        gravity_scale=self._gravity_scale, check_args=False, user_data=self.user_data)
        for fixture in self.fixtures:
            new_body.attach_fixture(copy(fixture))
        return new_body

    # -- helpers ---------------------------------------------------------------------------
    def _xf_mul(self, v):
        px, py, c, s = self._xf
        return (px + c * v[0] + (-s) * v[1], py + s * v[0] + c * v[1])

    def _check_unlocked(self):
        if self._world is not None and self._world.locked:
            raise LockedError("world is locked (stepping)")

    def _dirty(self):
        if self._world is not None:
            self._world._host_dirty = True

    # -- fixtures --------------------------------------------------------------------------
    def create_fixture(self, shape, friction=0.2, restitution=0.0, density=0.0,
                       category_bits=0x0001, mask_bits=0xFFFF, group_index=0,
                       sensor=False, user_data=None):
        self._check_unlocked()
        fixture = Fixture(shape, friction, restitution, density, category_bits,
                          mask_bits, group_index, sensor, user_data, body=self)
        self.attach_fixture(fixture)
        return fixture

    def attach_fixture(self, fixture):
        if fixture in self.fixtures:
            raise ValueError("Fixture is already attached to this body")
        fixture._attach_to_body(self)
        self.fixtures.append(fixture)
        if fixture._density > 0.0:
            self.reset_mass_data()
        if self._world is not None:
            self._world._fixture_attached(self, fixture)

    def destroy_fixture(self, fixture):
        """body.py:258-286: the fixture leaves the body with its proxy and its contacts; the mass is recomputed."""
        self._check_unlocked()
        if fixture._body is not self or fixture not in self.fixtures:
            raise ValueError("This fixture does not belong to this body")
        self.fixtures.remove(fixture)
        fixture._body = None
        self.reset_mass_data()
        if self._world is not None:
            self._world._fixture_detached(self, fixture)

    def create_circle_fixture(self, radius=0.2, position=(0, 0), **kwargs):
        return self.create_fixture(_shapes.Circle(radius, position), **kwargs)

    def create_polygon_fixture(self, vertices=None, box=None, **kwargs):
        return self.create_fixture(_shapes.Polygon(vertices, box), **kwargs)

    def create_edge_fixture(self, v1=None, v2=None, v0=None, v3=None, **kwargs):
        return self.create_fixture(_shapes.Edge(v1, v2, v0, v3), **kwargs)

    def create_loop_fixture(self, vertices, *args, **kwargs):
        return self.create_fixture(_shapes.Loop(vertices, *args), **kwargs)

    def create_fixtures(self, shape_list, fixture=None, friction=0.2, restitution=0.0,
                        density=0.0, category_bits=0x0001, mask_bits=0xFFFF,
                        group_index=0, sensor=False):
        if isinstance(shape_list, _shapes.Shape):
            shape_list = [shape_list]
        if fixture is not None:
            friction, restitution, density = fixture.friction, fixture.restitution, fixture.density
            category_bits, mask_bits = fixture.category_bits, fixture.mask_bits
            group_index, sensor = fixture.group_index, fixture.sensor
        for shape in shape_list:
            self.create_fixture(shape, friction, restitution, density, category_bits,
                                mask_bits, group_index, sensor)

    # -- mass ------------------------------------------------------------------------------
    def reset_mass_data(self):
        """Sum fixture mass data; centre the inertia (body.py:606-660, same operation order)."""
        self._mass = 0.0
        self._inv_mass = 0.0
        self._I = 0.0
        self._invI = 0.0
        self._local_center = Vec2()
        if self._type in (Body.STATIC, Body.KINEMATIC):
            self._c = Vec2(self._xf[0], self._xf[1])
            self._c0 = copy(self._c)
            return
        cx = cy = 0.0
        for fixture in self.fixtures:
            if fixture._density == 0.0:
                continue
            md = fixture.mass_data
            self._mass += md.mass
            cx, cy = cx + md.mass * md.center.x, cy + md.mass * md.center.y
            self._I += md.I
        if self._mass > 0.0:
            self._inv_mass = 1.0 / self._mass
            cx, cy = self._inv_mass * cx, self._inv_mass * cy
        else:
            self._mass = 1.0
            self._inv_mass = 1.0
        if self._I > 0.0 and not self._fixed_rotation:
            self._I -= self._mass * (cx ** 2 + cy ** 2)
            assert self._I > 0.0
            self._invI = 1.0 / self._I
        else:
            self._I = 0.0
            self._invI = 0.0
        old = self._c
        self._local_center = Vec2(cx, cy)
        self._c0 = Vec2(*self._xf_mul(self._local_center))
        self._c = copy(self._c0)
        dx, dy = self._c.x - old.x, self._c.y - old.y
        w = self._angular_velocity
        self._linear_velocity = Vec2(self._linear_velocity.x + (-float(w) * dy),
                                     self._linear_velocity.y + (float(w) * dx))
        self._dirty()

    # -- read-backs ------------------------------------------------------------------------
    @property
    def world(self):
        return self._world

    @property
    def position(self):
        return Vec2(self._xf[0], self._xf[1])

    @position.setter
    def position(self, pos):
        # body.py:420-423: the origin moves, the rotation stays
        self._check_unlocked()
        self._before_transform_update()
        pos = Vec2(*pos)
        self._xf = (pos.x, pos.y, self._xf[2], self._xf[3])
        self._transform_updated()

    @property
    def angle(self):
        return math.atan2(self._xf[3], self._xf[2])

    @angle.setter
    def angle(self, angle):
        # body.py:430-433
        self._check_unlocked()
        self._before_transform_update()
        angle = float(angle)
        self._xf = (self._xf[0], self._xf[1], math.cos(angle), math.sin(angle))
        self._transform_updated()

    @property
    def transform(self):
        return Transform(self._xf[:2], cos_sin=self._xf[2:])

    @transform.setter
    def transform(self, position_angle):
        """body.py:384-395: `(position, angle)`, or a transform as this facade hands them out (`(px, py, cos, sin)`)."""
        self._check_unlocked()
        self._before_transform_update()
        if len(position_angle) == 4:
            self._xf = tuple(float(v) for v in position_angle)
        else:
            pos, angle = Vec2(*position_angle[0]), float(position_angle[1])
            self._xf = (pos.x, pos.y, math.cos(angle), math.sin(angle))
        self._transform_updated()

    def _before_transform_update(self):
        if self._world is not None:
            self._world._before_host_move(self)

    def _transform_updated(self):
        """body.py:404-413: the sweep restarts at the new transform (the angle read back from the rotation, as there), every
        fixture's proxy is moved with no displacement and the world looks for new contacts (`World._body_moved_by_host`)."""
        self._c0 = Vec2(*self._xf_mul(self._local_center))
        self._c = copy(self._c0)
        self._a0 = self._a = math.atan2(self._xf[3], self._xf[2])
        self._dirty()
        if self._world is not None:
            self._world._body_moved_by_host(self)

    @property
    def world_center(self):
        return copy(self._c)

    @property
    def local_center(self):
        return copy(self._local_center)

    @property
    def linear_velocity(self):
        return copy(self._linear_velocity)

    @linear_velocity.setter
    def linear_velocity(self, v):
        if self._type == Body.STATIC:
            return
        v = Vec2(*v)
        if v.dot(v) > 0.0:
            self.awake = True
        self._linear_velocity = v
        self._dirty()

    @property
    def angular_velocity(self):
        return self._angular_velocity

    @angular_velocity.setter
    def angular_velocity(self, w):
        if self._type == Body.STATIC:
            return
        if w * w > 0.0:
            self.awake = True
        self._angular_velocity = float(w)
        self._dirty()

    @property
    def mass(self):
        return self._mass

    @property
    def inertia(self):
        return self._I + self._mass * self._local_center.length_squared

    @property
    def mass_data(self):
        # body.py:563-575
        lc = self._local_center
        return _shapes.MassData(mass=self._mass, center=(lc.x, lc.y), I=self._I + self._mass * (lc.x * lc.x + lc.y * lc.y))

    @mass_data.setter
    def mass_data(self, md):
        """body.py:577-604: override the fixtures' mass properties (dynamic bodies only); the centre of mass moves and the
        linear velocity follows it."""
        self._check_unlocked()
        if self._type != Body.DYNAMIC:
            return
        self._I = 0.0
        self._invI = 0.0
        self._mass = md.mass
        if self._mass <= 0.0:
            self._mass = 1.0
        self._inv_mass = 1.0 / self._mass
        center = Vec2(*md.center)
        if md.I > 0.0 and not self._fixed_rotation:
            self._I = md.I - self._mass * (center.x * center.x + center.y * center.y)
            assert self._I > 0.0
            self._invI = 1.0 / self._I
        old = self._c
        self._local_center = center
        self._c0 = Vec2(*self._xf_mul(self._local_center))
        self._c = copy(self._c0)
        dx, dy = self._c.x - old.x, self._c.y - old.y
        w = float(self._angular_velocity)
        self._linear_velocity = Vec2(self._linear_velocity.x + (-w * dy), self._linear_velocity.y + (w * dx))
        self._dirty()

    @property
    def type(self):
        return self._type

    @type.setter
    def type(self, _type):
        """body.py:742-767: mass data recomputed for the new type, a body that turns static stops, the body is woken, its forces
        are dropped and the contacts of its fixtures are flagged for re-filtering (two bodies of which none is dynamic may not
        collide, body.py:910-918)."""
        self._check_unlocked()
        if self._type == _type:
            return
        if _type not in (Body.STATIC, Body.DYNAMIC, Body.KINEMATIC):
            raise ValueError("Invalid body type")
        self._type = _type
        self.reset_mass_data()
        if _type == Body.STATIC:
            self._linear_velocity = Vec2()
            self._angular_velocity = 0.0
        self.awake = True
        self._force = Vec2()
        self._torque = 0.0
        for fixture in self.fixtures:
            fixture._refilter()
        self._dirty()

    static = property(lambda self: self._type == Body.STATIC)
    dynamic = property(lambda self: self._type == Body.DYNAMIC)
    kinematic = property(lambda self: self._type == Body.KINEMATIC)
    @property
    def active(self):
        return self._active

    @active.setter
    def active(self, active):
        # body.py:848-872 creates / destroys the proxies and destroys the contacts.  Before the first step that is a host
        # flag (the proxies are built from it); afterwards it would be a host mutation of the device contact list
        self._check_unlocked()
        if self._active == bool(active):
            return
        if self._world is not None and self._world._stepped:
            # host mutation of the device proxy rows and contact list: applied when the world is packed again before the next
            # step.  Activation gives every fixture a fresh proxy in the move buffer (the pair pass at the END of the next step
            # finds its contacts: no new-fixture flag, body.py:858-862); deactivation takes the proxies and ALL of the body's
            # contacts away (the reference, iterating over the list it removes from, leaves every other one behind as a zombie
            # without proxies, body.py:868-870 with contact_manager.py:170-172: identical for a body with at most one contact).
            self._active = bool(active)
            self._world._body_activation_changed(self)
            return
        self._active = bool(active)
        if self._world is not None:
            self._world._batch = None          # proxies are rebuilt from the host state
            self._world._host_dirty = True
            self._world._body_activation_changed(self)

    # plain per-body parameters the step reads (body.py:710-790, 880-890): set on the host, uploaded before the next step
    def _parameter(name, convert, after=None):
        def fget(self):
            return getattr(self, name)

        def fset(self, value):
            setattr(self, name, convert(value))
            if after is not None:
                after(self)
            self._dirty()
        return property(fget, fset)

    def _woken_unless_allowed_to_sleep(self):
        if not self._allow_sleep:
            self.awake = True

    bullet = _parameter("_bullet", bool)
    fixed_rotation = _parameter("_fixed_rotation", bool, lambda self: self.reset_mass_data())
    linear_damping = _parameter("_linear_damping", float)
    angular_damping = _parameter("_angular_damping", float)
    gravity_scale = _parameter("_gravity_scale", float)
    allow_sleep = _parameter("_allow_sleep", bool, _woken_unless_allowed_to_sleep)
    del _parameter, _woken_unless_allowed_to_sleep

    @property
    def awake(self):
        return self._awake

    @awake.setter
    def awake(self, awake):
        # body.py:800-812
        if awake:
            if not self._awake:
                self._sleep_time = 0.0
        else:
            self._sleep_time = 0.0
            self._linear_velocity = Vec2()
            self._angular_velocity = 0.0
            self._force = Vec2()
            self._torque = 0.0
        self._awake = bool(awake)
        self._dirty()

    @property
    def sleeping(self):
        return not self._awake

    @sleeping.setter
    def sleeping(self, sleeping):
        self.awake = not sleeping

    # -- forces (host-side inputs consumed by the next step; body.py:477-545) ---------------
    def apply_force(self, force, point):
        if self._type != Body.DYNAMIC:
            return
        if not self._awake:
            self.awake = True
        f = Vec2(*force)
        p = Vec2(*point)
        self._force = self._force + f
        self._torque += (p - self._c).cross(f)
        self._dirty()

    def apply_force_to_center(self, force):
        """body.py:493-503: a force through the centre of mass (no torque); wakes the body."""
        if self._type != Body.DYNAMIC:
            return
        if not self._awake:
            self.awake = True
        self._force = self._force + Vec2(*force)
        self._dirty()

    # -- frames (body.py:662-708): host arithmetic on the transform mirrored from the device after every step -----------
    def get_world_point(self, local_point):
        x, y, c, s = self._xf
        p = Vec2(*local_point)
        return Vec2(x + c * p.x - s * p.y, y + s * p.x + c * p.y)

    def get_world_vector(self, local_vector):
        x, y, c, s = self._xf
        v = Vec2(*local_vector)
        return Vec2(c * v.x - s * v.y, s * v.x + c * v.y)

    def get_local_point(self, world_point):
        x, y, c, s = self._xf
        p = Vec2(*world_point)
        dx, dy = p.x - x, p.y - y
        return Vec2(c * dx + s * dy, -s * dx + c * dy)

    def get_local_vector(self, world_vector):
        x, y, c, s = self._xf
        v = Vec2(*world_vector)
        return Vec2(c * v.x + s * v.y, -s * v.x + c * v.y)

    def get_linear_velocity_from_world_point(self, world_point):
        p = Vec2(*world_point)
        rx, ry = p.x - self._c.x, p.y - self._c.y
        w = self._angular_velocity
        return Vec2(self._linear_velocity.x - w * ry, self._linear_velocity.y + w * rx)

    def get_linear_velocity_from_local_point(self, local_point):
        return self.get_linear_velocity_from_world_point(self.get_world_point(local_point))

    @property
    def contact_list(self):
        """Contacts attached to this body (body.py:889-896), read back from the device contact list."""
        if self._world is None:
            return []
        return [c for c in self._world.contact_manager.contacts if c.body_a is self or c.body_b is self]

    def apply_torque(self, torque):
        if self._type != Body.DYNAMIC:
            return
        if not self._awake:
            self.awake = True
        self._torque += torque
        self._dirty()

    def apply_linear_impulse(self, impulse, point):
        if self._type != Body.DYNAMIC:
            return
        if not self._awake:
            self.awake = True
        imp = Vec2(*impulse)
        p = Vec2(*point)
        self._linear_velocity = self._linear_velocity + self._inv_mass * imp
        self._angular_velocity += self._invI * (p - self._c).cross(imp)
        self._dirty()

    def apply_angular_impulse(self, impulse):
        if self._type != Body.DYNAMIC:
            return
        if not self._awake:
            self.awake = True
        self._angular_velocity += self._invI * impulse
        self._dirty()

    def __repr__(self):
        return "Body(position=%s, angle=%g, type=%d, fixtures=%d)" % (
            self.position, self.angle, self._type, len(self.fixtures))
