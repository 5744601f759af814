"""Fixture: a shape plus material and filter data attached to a body.

Mirrors the reference's constructor and attributes (`pypybox2d/fixture.py:52-69`) so that
scene scripts run unchanged.  Broadphase proxies are not Python objects here: each fixture
becomes one row of the per-world proxy arrays when the world is packed (`schema.pack_world`).
"""
from copy import copy

from .common import EmptyFixtureError, is_valid_float


class Fixture(object):
    def __init__(self, shape=None, friction=0.2, restitution=0.0, density=0.0,
                 category_bits=0x0001, mask_bits=0xFFFF, group_index=0,
                 sensor=False, user_data=None, body=None):
        self._friction = friction
        self._restitution = restitution
        self._density = density
        self._sensor = bool(sensor)      # a sensor collects contact information but never generates a collision response (fixture.py:163-168)
        self._category_bits = category_bits
        self._mask_bits = mask_bits
        self._group_index = group_index
        self._shape = copy(shape)        # shapes are cloned on attach (fixture.py:67)
        self.user_data = user_data
        self._body = None
        self._index = -1                 # row in the packed proxy arrays
        self._attach_to_body(body)

    def _attach_to_body(self, body):
        if body is not None and self._shape is None:
            raise EmptyFixtureError("Cannot attach empty fixture to a body")
        self._body = body

    def __copy__(self):
        return Fixture(shape=self._shape, friction=self._friction, restitution=self._restitution,
                       density=self._density, category_bits=self._category_bits,
                       mask_bits=self._mask_bits, group_index=self._group_index,
                       sensor=self._sensor, user_data=self.user_data)

    def __repr__(self):
        return "Fixture(shape=%s, friction=%g, restitution=%g, density=%g)" % (
            self._shape, self._friction, self._restitution, self._density)

    shape = property(lambda self: self._shape)
    body = property(lambda self: self._body)

    def _dirty(self):
        if self._body is not None:
            self._body._dirty()

    @property
    def friction(self):
        return self._friction

    @friction.setter
    def friction(self, friction):
        # fixture.py:130-132; existing contacts keep the mixed value they were created with (contact.py:165-166)
        self._friction = friction
        self._dirty()

    @property
    def restitution(self):
        return self._restitution

    @restitution.setter
    def restitution(self, restitution):
        self._restitution = restitution      # fixture.py:142-144
        self._dirty()

    @property
    def density(self):
        return self._density

    @density.setter
    def density(self, density):
        # fixture.py:155-160: takes effect at the body's next reset_mass_data
        if not is_valid_float(density) or density < 0:
            raise ValueError("Invalid density")
        self._density = density

    @property
    def sensor(self):
        return self._sensor

    @sensor.setter
    def sensor(self, sensor):
        # fixture.py:169-174: the body is woken.  The reference's narrowphase reads the flag at every update (contact.py:229-239);
        # the device keeps it per contact from the contact's creation, so the world is packed again before the next step and the
        # flag of every carried contact is set from its two fixtures (World._rebuild_batch)
        if bool(sensor) != self._sensor:
            if self._body is not None:
                self._body.awake = True
            self._sensor = bool(sensor)
            self._dirty()
            world = self.world
            if world is not None and world._stepped:
                world._structure_dirty = True

    # collision filter (fixture.py:176-225): the fixture's contacts are flagged for re-filtering and its proxies touched
    def _filter_field(name):
        def fget(self):
            return getattr(self, name)

        def fset(self, value):
            setattr(self, name, int(value))
            self._refilter()
        return property(fget, fset)

    category_bits = _filter_field("_category_bits")
    mask_bits = _filter_field("_mask_bits")
    group_index = _filter_field("_group_index")
    del _filter_field

    def _refilter(self):
        # fixture.py:269-285
        if self._body is None:
            return
        world = self.world
        if world is not None:
            world._fixture_refiltered(self)

    def get_aabb(self, child_index=0):
        """The fixture's broadphase AABB as ((lx, ly), (ux, uy)) -- the swept box of the last synchronisation, possibly stale
        (fixture.py:242-248) -- read back from the device proxy rows."""
        world = self.world
        if world is None or world._batch is None:
            lx, ly, ux, uy = self._shape.compute_aabb(self._body._xf)
            return (lx, ly), (ux, uy)
        j = world._fixture_list().index(self)
        t = world._ensure_batch().download(["f_tight"], worlds=slice(0, 1))["f_tight"][0, j]
        return (float(t[0]), float(t[1])), (float(t[2]), float(t[3]))

    def test_point(self, p):
        """Is the world point inside this fixture's shape?  (fixture.py:228-233; evaluated on the device)"""
        world = self.world
        if world is None:
            raise ValueError("fixture is not attached to a world")
        batch = world._ensure_batch()
        return bool(batch.test_points([0], [world._fixture_list().index(self)], [(p[0], p[1])])[0])

    def ray_cast(self, p1, p2, max_fraction=1.0, child_index=0):
        """(hit, normal, fraction) of the segment against this fixture's shape (fixture.py:235-240; evaluated on the device)."""
        from .common import Vec2
        world = self.world
        if world is None:
            raise ValueError("fixture is not attached to a world")
        batch = world._ensure_batch()
        me = world._fixture_list().index(self)
        count, fix, hit = batch.ray_cast([0], [(p1[0], p1[1], p2[0], p2[1])], max_hits=max(len(world._fixture_list()), 1))
        for k in range(max(int(count[0]), 0)):
            if int(fix[0][k]) == me and float(hit[0][k][0]) <= max_fraction:
                return True, Vec2(float(hit[0][k][1]), float(hit[0][k][2])), float(hit[0][k][0])
        return False, None, 0.0

    @property
    def mass_data(self):
        return self._shape.compute_mass(self._density)

    @property
    def world(self):
        return None if self._body is None else self._body._world
