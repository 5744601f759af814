"""Structure-of-arrays layout of a batch of worlds (the data the C-ABI borrows).

Every array is world-major: `[W, cap, k]`, so one CTA streams its own world's rows with
unit stride.  The field set is the persistent state of the reference's objects
(SURVEY §8 a23 / Appendix B): `Body` (`pypybox2d/body.py:65-103`), `Sweep`
(`common.py:870-889`), `Fixture`/`FixtureProxy` (`fixture.py:31-69`), `Manifold`
(`contact_util.py:50-166`) and `Contact` flags (`contact.py:57-81`).

`FIELDS` is the single source of truth: `include/b2b.h` (`b2b_buffers`) and
the test-only checker under `oracle/` list the same names in the same order, and
`tests/test_abi.py` checks that through `b2b_field_names()` / `orc_field_names()`.
"""
import numpy as np

from . import shapes as _shapes

# ---- flag bits -------------------------------------------------------------------------------
BODY_TYPE_MASK = 0x3          # 0 static, 1 dynamic, 2 kinematic  (body.py:36-38)
BODY_AWAKE = 0x4
BODY_ACTIVE = 0x8
BODY_ALLOW_SLEEP = 0x10
BODY_FIXED_ROT = 0x20
BODY_BULLET = 0x40
BODY_ISLAND = 0x80

CONTACT_TOUCHING = 0x1        # contact.py:73
CONTACT_ENABLED = 0x2         # contact.py:75
CONTACT_FILTER = 0x4          # contact.py:77
CONTACT_ISLAND = 0x8          # contact.py:71
CONTACT_SENSOR = 0x10        # one of the two fixtures is a sensor (fixture.py:163-168): no manifold, touching = GJK overlap

WORLD_ALLOW_SLEEP = 0x1       # world.py:77
WORLD_WARM_STARTING = 0x2     # world.py:74
WORLD_NEW_FIXTURE = 0x4       # world.py:85
WORLD_AUTO_CLEAR = 0x8        # world.py:80
WORLD_GRAPH_VALID = 0x10      # device bookkeeping: last step's island lists / schedule are still valid
WORLD_CONTINUOUS = 0x20       # world.py:78 continuous_physics: World.solve_toi runs after the discrete solve
WORLD_TOI_DIRTY = 0x40        # device bookkeeping: some sweep.alpha0 (b_pos0[3]) is non-zero and is reset by the next World.solve_toi

STATUS_CONTACT_OVERFLOW = 0x1
STATUS_PAIR_OVERFLOW = 0x2
STATUS_NAN = 0x4
STATUS_ISLAND_OVERFLOW = 0x8
STATUS_TOI_STUCK = 0x10        # a time-of-impact push-back loop was cut off (csrc/b2b_toi.cuh TOI_MAX_PUSH_BACK)
STATUS_JOINT_OVERFLOW = 0x20   # more joints than the batch's joint capacity

# joints (pypybox2d/joints/*.py): j_def = (type, body_a, body_b, flags)
JOINT_REVOLUTE, JOINT_DISTANCE, JOINT_WELD, JOINT_PRISMATIC, JOINT_ROPE, JOINT_FRICTION, JOINT_WHEEL, JOINT_PULLEY, JOINT_MOUSE = 1, 2, 3, 4, 5, 6, 7, 8, 9
JOINT_COLLIDE_CONNECTED = 0x1  # joint.py:52
JOINT_LIMIT_ENABLED = 0x2      # revolute.py:85 / prismatic
JOINT_MOTOR_ENABLED = 0x4      # revolute.py:86 / prismatic
JOINT_ISLAND = 0x8             # Joint._island_flag (transient bookkeeping of the island walk)
JOINT_LIMIT_SHIFT = 8          # bits 8..9: _limit_state (joint.py:36: INACTIVE, AT_LOWER, AT_UPPER, EQUAL), carried over steps
JOINT_PAR = 12                 # reals per joint parameter row (see joints.py `_par`)
JOINT_TMP = 28                 # reals per joint of solver temporaries (ra, rb, effective masses ...; [20:24] the bodies' inverse masses,
                               # [24:28] the local anchors relative to the centres of mass), written by the step's joint set-up

MANIFOLD_CIRCLES, MANIFOLD_FACE_A, MANIFOLD_FACE_B = 0, 1, 2   # contact_util.py:141-143

NO_PROXY = 1.0e30             # fat / tight AABB (lower = +NO_PROXY, upper = -NO_PROXY) of a fixture without a proxy (inactive body)

SHAPE_DATA = 32               # reals per shape row: polygon = 8 verts + 8 normals (x,y)

# (name, scope, trailing shape, kind).  scope: W = per world, B = per body, F = per fixture/proxy,
# C = per contact, P = per new-pair slot, L = per island-list slot, S = shared shape table.
FIELDS = [
    # -- per world --
    ("w_count",   "W", (4,), "i32"),   # n_bodies, n_fixtures, n_contacts, n_pairs(last update_pairs)
    ("w_flags",   "W", (),   "u32"),
    ("w_status",  "W", (),   "u32"),
    ("w_gravity", "W", (2,), "real"),
    ("w_island",  "W", (4,), "i32"),   # n_islands, n_island_bodies, n_island_contacts, pos_iters_max
    # -- per body --
    ("b_xf",      "B", (4,), "real"),  # origin x, y, cos, sin            (Body._xf)
    ("b_pos",     "B", (4,), "real"),  # sweep.c x, y, sweep.a, sleep_time
    ("b_vel",     "B", (4,), "real"),  # v x, y, w, torque
    ("b_pos0",    "B", (4,), "real"),  # sweep.c0 x, y, sweep.a0, sweep.alpha0
    ("b_force",   "B", (2,), "real"),
    ("b_mass",    "B", (4,), "real"),  # inv_mass, inv_I, local_center x, y
    ("b_damp",    "B", (4,), "real"),  # linear_damping, angular_damping, gravity_scale, (unused)
    ("b_flags",   "B", (),   "u32"),
    ("b_island",  "B", (),   "i32"),   # Body._island_index of the last step (-1: not in an island)
    # -- per fixture == per broadphase proxy (all supported shapes have one child) --
    ("f_body",    "F", (),   "i32"),
    ("f_shape",   "F", (),   "i32"),   # row of the shared shape table
    ("f_mat",     "F", (2,), "real"),  # friction, restitution
    ("f_filter",  "F", (2,), "u32"),   # category | mask << 16 ; (group & 0xffff) | sensor << 16
    ("f_fat",     "F", (4,), "real"),  # tree-node (fat) AABB lower x, y, upper x, y
    ("f_tight",   "F", (4,), "real"),  # FixtureProxy.aabb (swept tight AABB)
    ("f_moved",   "F", (),   "u8"),    # in BroadPhase._move_buffer
    # -- per contact, creation order --
    ("c_fix",     "C", (2,), "i32"),
    ("c_flags",   "C", (),   "u32"),
    ("c_mat",     "C", (2,), "real"),  # mixed friction, restitution
    ("c_mtype",   "C", (),   "i32"),   # manifold.type | point_count << 8
    ("c_mf",      "C", (4,), "real"),  # local_normal x, y, local_point x, y
    ("c_pts",     "C", (4,), "real"),  # points[0].local_point x, y, points[1].local_point x, y
    ("c_imp",     "C", (4,), "real"),  # normal_impulse0, tangent_impulse0, normal_impulse1, tangent_impulse1
    ("c_id",      "C", (2,), "u32"),   # id tuple packed a | b<<8 | c<<16 | d<<24
    # -- observable scratch of the last step --
    ("p_pairs",   "P", (2,), "i32"),   # pairs yielded by the last update_pairs, in emission order
    ("l_body",    "L", (),   "i32"),   # island body lists, concatenated in island order
    ("l_contact", "C", (),   "i32"),   # island contact lists, concatenated in island order
    ("l_start",   "B", (2,), "i32"),   # per island: start in l_body, start in l_contact (ends: next row / w_island)
    # -- shared shape table --
    ("s_meta",    "S", (),   "i32"),   # type | vertex_count << 8 | has_v0 << 16 | has_v3 << 17 | loop_child << 18
    ("s_radius",  "S", (),   "real"),
    ("s_data",    "S", (SHAPE_DATA,), "real"),
    # -- joints, creation order == World.joints; appended so that the rows above keep their positions --
    ("w_joint",   "W", (4,), "i32"),   # n_joints, n_island_joints (last step), unused, unused
    ("j_def",     "J", (4,), "i32"),   # type, body_a, body_b, flags (JOINT_* bits | _limit_state << 8)
    ("j_par",     "J", (JOINT_PAR,), "real"),  # constructor parameters (joints.py `_par`)
    ("j_imp",     "J", (4,), "real"),  # accumulated impulses: _impulse x, y, z, _motor_impulse (distance: _impulse in [0])
    ("j_tmp",     "J", (JOINT_TMP,), "real"),  # solver temporaries of the current step
    ("j_adj",     "J", (2,), "i32"),   # Body.joints of every body as rows of joint indices, flattened (2 entries per joint)
    ("b_jadj",    "B", (2,), "i32"),   # this body's row in j_adj: start, count (body.py:75, world.py:270-273 append order)
    ("l_joint",   "J", (),   "i32"),   # island joint lists, concatenated in island order
    ("l_jstart",  "B", (),   "i32"),   # per island: start in l_joint
]

FIELD_NAMES = [f[0] for f in FIELDS]


class Caps(object):
    """Per-world capacities.  `nl` (island-list slots) and `npair` default from the others."""

    def __init__(self, nb, nf, nc, npair=None, nl=None, nj=0):
        self.nj = int(nj)           # joints per world (0: the batch has none; the J arrays keep one unused row)
        self.nb = int(nb)
        self.nf = int(nf)
        self.nc = int(nc)
        self.npair = int(npair if npair is not None else 2 * nc)
        self.nl = int(nl if nl is not None else nb + nc)

    def dim(self, scope, n_worlds, n_shapes):
        return {"W": (n_worlds,), "B": (n_worlds, self.nb), "F": (n_worlds, self.nf),
                "C": (n_worlds, self.nc), "P": (n_worlds, self.npair), "L": (n_worlds, self.nl),
                "J": (n_worlds, max(self.nj, 1)), "S": (n_shapes,)}[scope]

    def __repr__(self):
        return "Caps(nb=%d, nf=%d, nc=%d, npair=%d, nl=%d, nj=%d)" % (self.nb, self.nf, self.nc, self.npair, self.nl, self.nj)


def np_dtype(kind, real):
    return {"real": np.float64 if real == "f64" else np.float32, "i32": np.int32, "u32": np.uint32,
            "u8": np.uint8}[kind]


def alloc(caps, n_worlds, n_shapes, real="f64"):
    """Zero-initialised host arrays for a batch."""
    out = {}
    for name, scope, tail, kind in FIELDS:
        out[name] = np.zeros(caps.dim(scope, n_worlds, n_shapes) + tail, dtype=np_dtype(kind, real))
    out["b_island"][...] = -1
    return out


def pack_id(t):
    return (int(t[0]) & 0xff) | ((int(t[1]) & 0xff) << 8) | ((int(t[2]) & 0xff) << 16) | ((int(t[3]) & 0xff) << 24)


class ShapeTable(object):
    """De-duplicated shape rows shared by every world of a batch."""

    def __init__(self):
        self._rows = {}
        self.meta = []
        self.radius = []
        self.data = []

    def add(self, shape):
        key = shape._key()
        row = self._rows.get(key)
        if row is not None:
            return row
        data = [0.0] * SHAPE_DATA
        if shape._type == _shapes.CIRCLE:
            meta = _shapes.CIRCLE | (1 << 8)
            data[0], data[1] = shape.position.x, shape.position.y
        elif shape._type == _shapes.POLYGON:
            n = len(shape._vertices)
            meta = _shapes.POLYGON | (n << 8)
            for i, (v, nrm) in enumerate(zip(shape._vertices, shape._normals)):
                data[2 * i], data[2 * i + 1] = v.x, v.y
                data[16 + 2 * i], data[16 + 2 * i + 1] = nrm.x, nrm.y
        elif shape._type == _shapes.EDGE:
            meta = _shapes.EDGE | (2 << 8)
            data[0], data[1] = shape._vertex1.x, shape._vertex1.y
            data[2], data[3] = shape._vertex2.x, shape._vertex2.y
            if shape._vertex0 is not None:
                meta |= 1 << 16
                data[4], data[5] = shape._vertex0.x, shape._vertex0.y
            if shape._vertex3 is not None:
                meta |= 1 << 17
                data[6], data[7] = shape._vertex3.x, shape._vertex3.y
            if getattr(shape, "_loop_child", False):
                meta |= 1 << 18              # a Loop's child edge: AABB without the radius margin (shapes.py:822-837)
        else:
            raise NotImplementedError("unsupported shape type")
        row = len(self.meta)
        self._rows[key] = row
        self.meta.append(meta)
        self.radius.append(float(shape.radius))
        self.data.append(data)
        return row

    def __len__(self):
        return len(self.meta)

    def arrays(self, real="f64"):
        rd = np_dtype("real", real)
        n = max(len(self.meta), 1)
        meta = np.zeros((n,), np.int32)
        radius = np.zeros((n,), rd)
        data = np.zeros((n, SHAPE_DATA), rd)
        if self.meta:
            meta[:len(self.meta)] = self.meta
            radius[:len(self.meta)] = self.radius
            data[:len(self.meta)] = np.asarray(self.data, dtype=np.float64)
        return meta, radius, data


def world_proxies(world):
    """[(fixture, child index)] in world.bodies x body.fixtures x children order: one row of the F arrays each (a fixture has
    one broadphase proxy per child of its shape: fixture.py:250-261; only Loop shapes have more than one child)."""
    return [(f, c) for b in world.bodies for f in b.fixtures for c in range(f._shape.child_count)]


def proxy_shape(fixture, child):
    """The shape a proxy collides as: the fixture's shape, or the child edge of its Loop (contact.py:363-373)."""
    s = fixture._shape
    return s.child_edge(child) if s._type == _shapes.LOOP else s


def add_shapes(table, world):
    """Every proxy shape of `world` into the shape table (de-duplicated)."""
    for f, c in world_proxies(world):
        table.add(proxy_shape(f, c))


def body_flags(body):
    f = body._type & BODY_TYPE_MASK
    if body._awake:
        f |= BODY_AWAKE
    if body._active:
        f |= BODY_ACTIVE
    if body._allow_sleep:
        f |= BODY_ALLOW_SLEEP
    if body._fixed_rotation:
        f |= BODY_FIXED_ROT
    if body._bullet:
        f |= BODY_BULLET
    return f


def pack_world(world, arrays, w, table):
    """Write the host `World` facade into row `w` of `arrays` (fp64 values, cast on store).

    Proxy rows are created in `world.bodies` x `body.fixtures` order, which is the order of the
    reference's move buffer and of the pinned pair sort key.  Fat AABBs
    are tight +- AABB_EXTENSION as `DynamicTree.create_proxy` makes them (dynamictree.py:122-126).
    """
    from .settings import AABB_EXTENSION
    nb = len(world.bodies)
    proxies = world_proxies(world)
    nf = len(proxies)
    A = arrays
    if nb > A["b_flags"].shape[1] or nf > A["f_body"].shape[1]:
        raise ValueError("world exceeds batch capacity")
    A["w_count"][w] = (nb, nf, 0, 0)
    flags = WORLD_NEW_FIXTURE
    if world._allow_sleep:
        flags |= WORLD_ALLOW_SLEEP
    if world._warm_starting:
        flags |= WORLD_WARM_STARTING
    if world._auto_clear_forces:
        flags |= WORLD_AUTO_CLEAR
    if getattr(world, "_continuous_physics", False):
        flags |= WORLD_CONTINUOUS
    A["w_flags"][w] = flags
    A["w_status"][w] = 0
    A["w_gravity"][w] = (world._gravity.x, world._gravity.y)
    index = {}
    for i, b in enumerate(world.bodies):
        index[id(b)] = i
        A["b_xf"][w, i] = b._xf
        A["b_pos"][w, i] = (b._c.x, b._c.y, b._a, b._sleep_time)
        A["b_vel"][w, i] = (b._linear_velocity.x, b._linear_velocity.y, b._angular_velocity, b._torque)
        A["b_pos0"][w, i] = (b._c0.x, b._c0.y, b._a0, 0.0)
        A["b_force"][w, i] = (b._force.x, b._force.y)
        A["b_mass"][w, i] = (b._inv_mass, b._invI, b._local_center.x, b._local_center.y)
        A["b_damp"][w, i] = (b._linear_damping, b._angular_damping, b._gravity_scale, 0.0)
        A["b_flags"][w, i] = body_flags(b)
        A["b_island"][w, i] = -1
    for j, (f, child) in enumerate(proxies):
        if child == 0:
            f._index = j
        body = f._body
        shape = proxy_shape(f, child)
        A["f_body"][w, j] = index[id(body)]
        A["f_shape"][w, j] = table.add(shape)
        A["f_mat"][w, j] = (f._friction, f._restitution)
        A["f_filter"][w, j] = ((f._category_bits & 0xffff) | ((f._mask_bits & 0xffff) << 16),
                               (f._group_index & 0xffff) | ((1 if f._sensor else 0) << 16))
        if not body._active:
            # the fixtures of an inactive body have no proxies (body.py:288-297, 848-872): an inverted box overlaps nothing, the
            # row never enters the move buffer, so no pair and no contact can ever name it
            A["f_tight"][w, j] = A["f_fat"][w, j] = (NO_PROXY, NO_PROXY, -NO_PROXY, -NO_PROXY)
            A["f_moved"][w, j] = 0
            continue
        lx, ly, ux, uy = shape.compute_aabb(body._xf)
        A["f_tight"][w, j] = (lx, ly, ux, uy)
        A["f_fat"][w, j] = (lx - AABB_EXTENSION, ly - AABB_EXTENSION, ux + AABB_EXTENSION, uy + AABB_EXTENSION)
        A["f_moved"][w, j] = 1
    from .joints import pack_joints
    pack_joints(world, A, w, index)
    return nb, nf
