"""Harness around the LIVE reference (pure-Python pypybox2d) -- TEST INFRASTRUCTURE ONLY.

Runs only where the reference tree exists (this container: /root/reference, read-only; never on
the GPU box).  It (1) imports the unmodified reference, (2) pins the pair-emission order of
`BroadPhase.update_pairs` (the reference iterates a `set` of objects hashed by address, so its
contact creation order differs between processes -- SURVEY §0 finding 1; any fixed order is a
legal execution), and (3) extracts the live objects' state into the schema arrays
(pypybox2d_b200/schema.py) so the C oracle and the CUDA path can be compared field by field.
"""
import os
import sys

import numpy as np

REF_PATH = os.environ.get("B2B_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_PATH, "pypybox2d"))


_b2 = None


def load():
    """Import the reference (pure Python: no `_common` extension is built in the read-only tree)."""
    global _b2
    if _b2 is not None:
        return _b2
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_PATH)
    import contextlib
    import io
    sys.path.insert(0, REF_PATH)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import pypybox2d as b2
    finally:
        sys.path.remove(REF_PATH)
    from pypybox2d import broadphase as _bp

    def update_pairs(self, callback=None):
        # broadphase.py:83-109 with the `set` replaced by an insertion-ordered dict and a final sort
        # by (proxy index a, proxy index b); proxy index = world.bodies x body.fixtures order.
        pairs = {}
        for proxy in self._move_buffer:
            for hit in self._tree.query(proxy.aabb):
                if hit is proxy:
                    continue
                if (id(hit), id(proxy)) not in pairs:
                    pairs[(id(proxy), id(hit))] = (proxy, hit)
        self._move_buffer = []
        out = sorted(pairs.values(), key=lambda p: (_pid(p[0]), _pid(p[1])))
        rec = getattr(self, "_pair_log", None)
        if rec is not None:
            rec.append([(_pid(a), _pid(b)) for a, b in out])
        for a, b in out:
            if callback:
                callback(a.user_data, b.user_data)
            else:
                yield (a.user_data, b.user_data)

    _bp.BroadPhase.update_pairs = update_pairs
    _b2 = b2
    return b2


_INDEX = {}


def _pid(node):
    p = node.user_data
    return _INDEX[(id(p.fixture), p.child_index)]


def index_fixtures(world):
    """Assign proxy indices in world.bodies x body.fixtures order (the pinned sort key)."""
    k = 0
    for body in world.bodies:
        for f in body.fixtures:
            for c in range(f._shape.child_count):
                _INDEX[(id(f), c)] = k
                k += 1
    return k


def pack_id(t):
    return (int(t[0]) & 0xff) | ((int(t[1]) & 0xff) << 8) | ((int(t[2]) & 0xff) << 16) | ((int(t[3]) & 0xff) << 24)


def extract(world, arrays, w, table, schema):
    b2 = load()
    """Copy the reference world's live state into row `w` of fp64 schema arrays.

    Sources are the attributes listed in SURVEY Appendix C.  `table` is a schema.ShapeTable fed with
    adapter objects so shape rows are de-duplicated exactly as the product's packer does.
    """
    A = arrays
    bodies = world.bodies
    bidx = {id(b): i for i, b in enumerate(bodies)}
    # one row per broadphase proxy: (fixture, child index) in world.bodies x body.fixtures x children order (Loop shapes have
    # one proxy per child edge, fixture.py:250-261)
    proxies = [(f, c) for b in bodies for f in b.fixtures for c in range(f._shape.child_count)]
    fixtures = [f for f, c in proxies]
    fidx = {(id(f), c): j for j, (f, c) in enumerate(proxies)}
    contacts = world.contact_manager.contacts
    A["w_count"][w] = (len(bodies), len(fixtures), len(contacts), A["w_count"][w][3])
    flags = 0
    if world._allow_sleep:
        flags |= schema.WORLD_ALLOW_SLEEP
    if world._warm_starting:
        flags |= schema.WORLD_WARM_STARTING
    if world._new_fixture:
        flags |= schema.WORLD_NEW_FIXTURE
    if world._auto_clear_forces:
        flags |= schema.WORLD_AUTO_CLEAR
    if world._continuous_physics:
        flags |= schema.WORLD_CONTINUOUS
    A["w_flags"][w] = flags
    A["w_gravity"][w] = (world._gravity.x, world._gravity.y)
    for i, b in enumerate(bodies):
        xf = b._xf
        A["b_xf"][w, i] = (xf._position.x, xf._position.y, xf._rotation._col1.x, xf._rotation._col1.y)
        s = b._sweep
        A["b_pos"][w, i] = (s.c.x, s.c.y, s.a, b._sleep_time)
        A["b_vel"][w, i] = (b._linear_velocity.x, b._linear_velocity.y, b._angular_velocity, b._torque)
        A["b_pos0"][w, i] = (s.c0.x, s.c0.y, s.a0, s.alpha0)
        A["b_force"][w, i] = (b._force.x, b._force.y)
        A["b_mass"][w, i] = (b._inv_mass, b._invI, s.local_center.x, s.local_center.y)
        A["b_damp"][w, i] = (b._linear_damping, b._angular_damping, b._gravity_scale, 0.0)
        f = b._type & 3
        if b._awake:
            f |= schema.BODY_AWAKE
        if b._active:
            f |= schema.BODY_ACTIVE
        if b._allow_sleep:
            f |= schema.BODY_ALLOW_SLEEP
        if b._fixed_rotation:
            f |= schema.BODY_FIXED_ROT
        if b._bullet:
            f |= schema.BODY_BULLET
        if b._island_flag:
            f |= schema.BODY_ISLAND
        A["b_flags"][w, i] = f
    moved = set(id(n) for n in world.contact_manager.broadphase._move_buffer)
    for j, (fx, child) in enumerate(proxies):
        A["f_body"][w, j] = bidx[id(fx._body)]
        A["f_shape"][w, j] = table.add(_ShapeAdapter(fx._shape.child_edge(child), loop_child=True) if isinstance(fx._shape, b2.Loop)
                                       else _ShapeAdapter(fx._shape))
        A["f_mat"][w, j] = (fx._friction, fx._restitution)
        A["f_filter"][w, j] = ((fx._category_bits & 0xffff) | ((fx._mask_bits & 0xffff) << 16),
                               (fx._group_index & 0xffff) | ((1 if fx._sensor else 0) << 16))
        proxy = fx._proxies[child]
        node = proxy.proxy_ref
        A["f_fat"][w, j] = (node.aabb.lower_bound.x, node.aabb.lower_bound.y,
                            node.aabb.upper_bound.x, node.aabb.upper_bound.y)
        A["f_tight"][w, j] = (proxy.aabb.lower_bound.x, proxy.aabb.lower_bound.y,
                              proxy.aabb.upper_bound.x, proxy.aabb.upper_bound.y)
        A["f_moved"][w, j] = 1 if id(node) in moved else 0
    for k, c in enumerate(contacts):
        A["c_fix"][w, k] = (fidx[(id(c._fixture_a), c._index_a)], fidx[(id(c._fixture_b), c._index_b)])
        cf = 0
        if c._touching_flag:
            cf |= schema.CONTACT_TOUCHING
        if c._enabled_flag:
            cf |= schema.CONTACT_ENABLED
        if c._filter_flag:
            cf |= schema.CONTACT_FILTER
        if c._island_flag:
            cf |= schema.CONTACT_ISLAND
        if c._fixture_a.sensor or c._fixture_b.sensor:
            cf |= schema.CONTACT_SENSOR
        A["c_flags"][w, k] = cf
        A["c_mat"][w, k] = (c._friction, c._restitution)
        m = c._manifold
        A["c_mtype"][w, k] = m.type | (m.point_count << 8)
        A["c_mf"][w, k] = (m.local_normal.x, m.local_normal.y, m.local_point.x, m.local_point.y)
        p0, p1 = m.points[0], m.points[1]
        A["c_pts"][w, k] = (p0.local_point.x, p0.local_point.y, p1.local_point.x, p1.local_point.y)
        A["c_imp"][w, k] = (p0.normal_impulse, p0.tangent_impulse, p1.normal_impulse, p1.tangent_impulse)
        A["c_id"][w, k] = (pack_id(p0.id), pack_id(p1.id))
    extract_joints(world, A, w, bidx, schema)
    return len(bodies), len(fixtures), len(contacts)


def extract_joints(world, A, w, bidx, schema):
    """World.joints / Body.joints of the live reference into the joint rows (layout: pypybox2d_b200/joints.py)."""
    b2 = load()
    from pypybox2d import joints as _j
    joints = world.joints
    if len(joints) > A["j_def"].shape[1] and joints:
        raise ValueError("joint capacity")
    A["w_joint"][w][0] = len(joints)
    jidx = {id(j): k for k, j in enumerate(joints)}
    for k, j in enumerate(joints):
        flags = 0
        if j._collide_connected:
            flags |= schema.JOINT_COLLIDE_CONNECTED
        if getattr(j, "_limit_enabled", False):
            flags |= schema.JOINT_LIMIT_ENABLED
        if getattr(j, "_motor_enabled", False):
            flags |= schema.JOINT_MOTOR_ENABLED
        flags |= int(getattr(j, "_limit_state", 0)) << schema.JOINT_LIMIT_SHIFT
        lb = j._local_anchor_b
        la = getattr(j, "_local_anchor_a", None) or b2.Vec2()
        par = [la.x, la.y, lb.x, lb.y]
        if isinstance(j, _j.RevoluteJoint):
            t = schema.JOINT_REVOLUTE
            par += [j._reference_angle, j._lower_angle, j._upper_angle, j._max_motor_torque, j._motor_speed]
            imp = (j._impulse.x, j._impulse.y, j._impulse.z, j._motor_impulse)
        elif isinstance(j, _j.DistanceJoint):
            t = schema.JOINT_DISTANCE
            par += [j._length, j._frequency, j._damping_ratio]
            imp = (j._impulse, 0.0, 0.0, 0.0)
        elif isinstance(j, _j.WeldJoint):
            t = schema.JOINT_WELD
            par += [j._reference_angle]
            imp = (j._impulse.x, j._impulse.y, j._impulse.z, 0.0)
        elif isinstance(j, _j.PrismaticJoint):
            t = schema.JOINT_PRISMATIC
            par += [j._reference_angle, j._lower_limit, j._upper_limit, j._max_motor_force, j._motor_speed,
                    j._local_x_axis.x, j._local_x_axis.y]
            imp = (j._impulse.x, j._impulse.y, j._impulse.z, j._motor_impulse)
        elif isinstance(j, _j.RopeJoint):
            t = schema.JOINT_ROPE
            par += [j._max_length]
            imp = (j._impulse, 0.0, 0.0, 0.0)
            flags |= int(j._state) << schema.JOINT_LIMIT_SHIFT
        elif isinstance(j, _j.FrictionJoint):
            t = schema.JOINT_FRICTION
            par += [j._max_force, j._max_torque]
            imp = (j._linear_impulse.x, j._linear_impulse.y, j._angular_impulse, 0.0)
        elif isinstance(j, _j.WheelJoint):
            t = schema.JOINT_WHEEL
            par += [j._local_x_axis.x, j._local_x_axis.y, j._frequency, j._damping_ratio, j._max_motor_torque, j._motor_speed]
            imp = (j._impulse, j._spring_impulse, j._motor_impulse, 0.0)
        elif isinstance(j, _j.PulleyJoint):
            t = schema.JOINT_PULLEY
            par += [j._ground_anchor_a.x, j._ground_anchor_a.y, j._ground_anchor_b.x, j._ground_anchor_b.y, j._ratio, j._constant]
            imp = (j._impulse, 0.0, 0.0, 0.0)
        elif isinstance(j, _j.MouseJoint):
            t = schema.JOINT_MOUSE
            par += [j._target.x, j._target.y, j._max_force, j._frequency, j._damping_ratio, j._body_b.mass]
            imp = (j._impulse.x, j._impulse.y, 0.0, 0.0)
        else:
            raise NotImplementedError(type(j))
        A["j_def"][w, k] = (t, bidx[id(j._body_a)] if j._body_a is not None else -1, bidx[id(j._body_b)], flags)
        A["j_par"][w, k] = 0.0
        A["j_par"][w, k, :len(par)] = par
        A["j_imp"][w, k] = imp
    pos = 0
    flat = A["j_adj"][w].reshape(-1)
    for i, b in enumerate(world.bodies):
        row = [jidx[id(j)] for j in b.joints]
        A["b_jadj"][w, i] = (pos, len(row))
        flat[pos:pos + len(row)] = row
        pos += len(row)


class _ShapeAdapter(object):
    """Presents a reference shape with the attribute names schema.ShapeTable.add reads."""

    def __init__(self, s, loop_child=False):
        b2 = load()
        self._loop_child = loop_child
        self.radius = s.radius
        if isinstance(s, b2.Circle):
            self._type = 0
            self.position = s.position
        elif isinstance(s, b2.Polygon):
            self._type = 1
            self._vertices = s._vertices
            self._normals = s._normals
        elif isinstance(s, b2.Edge):
            self._type = 3
            self._vertex1, self._vertex2 = s._vertex1, s._vertex2
            self._vertex0, self._vertex3 = s._vertex0, s._vertex3
        else:
            raise NotImplementedError(type(s))

    def _key(self):
        if self._type == 0:
            return (0, float(self.radius), self.position.x, self.position.y)
        if self._type == 1:
            return (1, float(self.radius)) + tuple(c for v in self._vertices for c in (v.x, v.y))
        v0, v3 = self._vertex0, self._vertex3
        return (3, float(self.radius), self._vertex1.x, self._vertex1.y, self._vertex2.x, self._vertex2.y,
                None if v0 is None else (v0.x, v0.y), None if v3 is None else (v3.x, v3.y), bool(self._loop_child))


class Recorder(object):
    """Wraps Island.solve and update_pairs of one reference world to log graded integer outputs."""

    def __init__(self, world):
        b2 = load()
        self.world = world
        self.islands = []      # per step: list of (body index list, contact index list)
        self.island_joints = []  # per step: list of joint index lists
        self.pairs = []        # per update_pairs call: list of (a, b)
        world.contact_manager.broadphase._pair_log = self.pairs
        from pypybox2d import island as _island
        rec = self
        if not hasattr(_island.Island, "_b2b_orig_solve"):
            _island.Island._b2b_orig_solve = _island.Island.solve

            def solve(self_i, step, gravity, allow_sleep):
                r = getattr(self_i, "_b2b_rec", None) or _ACTIVE.get("rec")
                if r is not None:
                    bi = {id(b): i for i, b in enumerate(r.world.bodies)}
                    ci = {id(c): i for i, c in enumerate(r.world.contact_manager.contacts)}
                    r.islands.append(([bi[id(b)] for b in self_i.bodies], [ci[id(c)] for c in self_i.contacts]))
                    ji = {id(j): i for i, j in enumerate(r.world.joints)}
                    r.island_joints.append([ji[id(j)] for j in self_i.joints])
                return _island.Island._b2b_orig_solve(self_i, step, gravity, allow_sleep)

            _island.Island.solve = solve
        _ACTIVE["rec"] = rec

    def begin_step(self):
        del self.islands[:]
        del self.island_joints[:]
        del self.pairs[:]


_ACTIVE = {}
