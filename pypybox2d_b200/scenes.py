"""Scene definitions of the five BASELINE.json configs.

Every builder takes the API module as its first argument, so the SAME script builds the scene
in the reference (`import pypybox2d as b2`, used by oracle/gen_golden.py) and in this package
(`import pypybox2d_b200 as b2`): the drop-in claim is exercised by construction.

World 0 of a config is the exact scripted scene.  World w > 0 jitters the initial position of
every dynamic body by `amp * U(-1, 1)` in x and y (SURVEY §8d "World diversity"), with U taken
from a counter-based hash of (seed, config, world, body, axis) so that a world's scene depends
only on its global index -- never on batch size, shard or GPU count (parity tier T4).
"""
import math

import numpy as np

SEED = 20261017
JITTER_AMP = 1e-2

_M64 = (1 << 64) - 1


def _mix64_np(x):
    """splitmix64 finaliser on uint64 arrays."""
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def u01(config, world, index, stream=0, seed=SEED):
    """Uniform [0,1) doubles; `world` and `index` broadcast.  53-bit mantissa from the hash."""
    world = np.asarray(world, dtype=np.uint64)
    index = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        k = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(config) * np.uint64(0xD1B54A32D192ED03)
             + np.uint64(stream) * np.uint64(0x8CB92BA72F3D8DD7))
        h = _mix64_np(_mix64_np(world + k) + index * np.uint64(0x9E3779B97F4A7C15))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def u01_torch(config, world, index, stream=0, seed=SEED):
    """Same numbers as `u01`, computed with torch int64 wrap-around arithmetic on any device."""
    import torch

    def i64(v):
        v &= _M64
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(x, s):
        return (x >> s) & ((1 << (64 - s)) - 1)

    def mix(x):
        x = (x ^ lsr(x, 30)) * i64(0xBF58476D1CE4E5B9)
        x = (x ^ lsr(x, 27)) * i64(0x94D049BB133111EB)
        return x ^ lsr(x, 31)

    k = i64(seed * 0x9E3779B97F4A7C15 + config * 0xD1B54A32D192ED03 + stream * 0x8CB92BA72F3D8DD7)
    world = world.to(torch.int64)
    index = index.to(torch.int64)
    h = mix(mix(world + k) + index * i64(0x9E3779B97F4A7C15))
    return lsr(h, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def jitter(config, world, n_bodies, amp=JITTER_AMP):
    """[n_bodies, 2] position offsets of world `world` (zeros for world 0), indexed by the body's position in
    `world.bodies` (hash index 2 * body + axis) -- the same numbers `BatchedWorld.replicate` draws on the device."""
    if world == 0 or amp == 0.0:
        return np.zeros((n_bodies, 2))
    idx = np.arange(2 * n_bodies, dtype=np.uint64)
    u = u01(config, world, idx).reshape(n_bodies, 2)
    return amp * (2.0 * u - 1.0)


def _kw(world_kw):
    """World keywords of a scene: continuous collision is OFF unless the caller asks for it (the parity setting)."""
    kw = dict(world_kw)
    kw.setdefault("continuous_physics", False)
    return kw


# ---- config 1 / 3: examples/pyramid-bench.py:32-53 -------------------------------------------
def pyramid(b2, world=0, rows=20, config=1, **world_kw):
    w = b2.World((0, -10), True, **_kw(world_kw))
    w.create_static_body(shapes=b2.Edge((-40, 0), (40, 0)))
    fixture = b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=5.0)
    n_dyn = rows * (rows + 1) // 2
    jit = jitter(config, world, n_dyn + 1)[1:]      # body 0 is the ground
    x = (-7.0, 0.75)
    k = 0
    for i in range(rows):
        y = x
        for j in range(i, rows):
            w.create_dynamic_body(position=(y[0] + jit[k, 0], y[1] + jit[k, 1]), fixtures=fixture)
            k += 1
            y = (y[0] + 1.125, y[1] + 0.0)
        x = (x[0] + 0.5625, x[1] + 1.25)
    return w


# ---- config 2(i): examples/hello-bench.py:21-36 ---------------------------------------------
def hello(b2, world=0, config=2, **world_kw):
    w = b2.World((0, -10), True, **_kw(world_kw))
    ground = w.create_static_body(position=(0, -10))
    ground.create_polygon_fixture(box=(50, 10))
    jit = jitter(config, world, 2)[1:]
    body = w.create_dynamic_body(position=(0 + jit[0, 0], 4 + jit[0, 1]))
    body.create_circle_fixture(1.0, density=1.0)
    return w


# ---- config 2(ii): examples/verticalstack-bench.py:33-75 ------------------------------------
def vertical_stack(b2, world=0, columns=5, rows=16, config=2, **world_kw):
    w = b2.World(**_kw(world_kw))
    w.create_static_body(shapes=[b2.Edge((-40, 0), (40, 0)), b2.Edge((20, 0), (20, 20))])
    box = b2.Fixture(shape=b2.Polygon(box=(0.5, 0.5)), density=1, friction=0.3)
    circle = b2.Fixture(shape=b2.Circle(0.5), density=1, friction=0.3)
    jit = jitter(config, world, 2 * columns * rows + 1)[1:]
    k = 0
    for j in range(columns):
        for i in range(rows):
            w.create_dynamic_body(fixtures=box, position=(-10 + 2.5 * j + jit[k, 0], 0.752 + 1.54 * i + jit[k, 1]))
            k += 1
            w.create_dynamic_body(fixtures=circle, position=(8 + 2.5 * j + jit[k, 0], 0.752 + 1.54 * i + jit[k, 1]))
            k += 1
    return w


# ---- config 4: mixed circles + convex polygons in a box of edges (testbed/confined.py:35-41) --
def mixed(b2, world=0, side=16, config=4, **world_kw):
    w = b2.World((0, -10), True, **_kw(world_kw))
    w.create_static_body(shapes=[b2.Edge((-10, 0), (10, 0)), b2.Edge((-10, 0), (-10, 20)),
                                 b2.Edge((10, 0), (10, 20)), b2.Edge((-10, 20), (10, 20))])
    n = side * side
    jit = jitter(config, world, n + 1)[1:]
    idx = np.arange(n, dtype=np.uint64)
    # shape draws come from their own streams and also vary with the world index (world 0 included)
    u_r = u01(config, world, idx, stream=1)
    u_m = u01(config, world, idx, stream=2)
    u_a = u01(config, world, idx, stream=3)
    for k in range(n):
        i, j = divmod(k, side)
        pos = (-9.0 + 1.2 * j + jit[k, 0], 0.8 + 1.2 * i + jit[k, 1])
        if k % 2 == 0:
            shape = b2.Circle(0.25 + 0.25 * float(u_r[k]))
        else:
            m = 3 + int(float(u_m[k]) * 6.0)
            radius = 0.3 + 0.3 * float(u_r[k])
            a0 = 2.0 * math.pi * float(u_a[k])
            shape = b2.Polygon(vertices=[(radius * math.cos(a0 + 2.0 * math.pi * q / m),
                                          radius * math.sin(a0 + 2.0 * math.pi * q / m)) for q in range(m)])
        w.create_dynamic_body(position=pos, fixtures=b2.Fixture(shape=shape, density=1, friction=0.3))
    return w


# ---- config 5: deep vertical stacks, 10 columns x 20 boxes -----------------------------------
def deep_stacks(b2, world=0, columns=10, rows=20, config=5, **world_kw):
    w = b2.World((0, -10), True, **_kw(world_kw))
    w.create_static_body(shapes=b2.Edge((-40, 0), (40, 0)))
    box = b2.Fixture(shape=b2.Polygon(box=(0.5, 0.5)), density=1, friction=0.3)
    jit = jitter(config, world, columns * rows + 1)[1:]
    k = 0
    for j in range(columns):
        for i in range(rows):
            w.create_dynamic_body(fixtures=box, position=(-12 + 2.5 * j + jit[k, 0], 0.752 + 1.54 * i + jit[k, 1]))
            k += 1
    return w


# ---- feature coverage scene (not a BASELINE config): every body / fixture option the step path reads ------------
def features(b2, world=0, config=6, **world_kw):
    """Edge chain with ghost vertices (EPCollider / edge-circle adjacency), a kinematic platform, restitution,
    damping, gravity scale, fixed rotation, a two-fixture body, group and category/mask filtering, circles and
    3..6-gons."""
    w = b2.World((0, -10), True, **_kw(world_kw))
    chain = [(-20.0, 6.0), (-12.0, 0.0), (-4.0, 0.0), (0.0, 2.5), (4.0, 0.0), (12.0, 0.0), (20.0, 6.0)]
    edges = []
    for i in range(len(chain) - 1):
        v0 = chain[i - 1] if i > 0 else None
        v3 = chain[i + 2] if i + 2 < len(chain) else None
        edges.append(b2.Edge(chain[i], chain[i + 1], v0, v3))
    w.create_static_body(shapes=edges, shape_fixture=b2.Fixture(b2.Circle(1.0), friction=0.6, restitution=0.1))
    w.create_kinematic_body(position=(-8.0, 3.0), linear_velocity=(0.8, 0.0), angular_velocity=0.1,
                            fixtures=b2.Fixture(b2.Polygon(box=(2.0, 0.25)), friction=0.5))
    n = 36
    jit = jitter(config, world, n + 2)[2:]
    for k in range(n):
        i, j = divmod(k, 6)
        pos = (-7.5 + 3.0 * j + jit[k, 0], 4.0 + 1.6 * i + jit[k, 1])
        kind = k % 6
        kw = {}
        if kind == 0:
            fx = [b2.Fixture(b2.Circle(0.4), density=1.0, friction=0.3, restitution=0.6)]
        elif kind == 1:
            fx = [b2.Fixture(b2.Polygon(box=(0.5, 0.35)), density=2.0, friction=0.4)]
            kw = dict(linear_damping=0.2, angular_damping=0.4)
        elif kind == 2:
            m = 3 + (k // 6) % 4
            fx = [b2.Fixture(b2.Polygon(vertices=[(0.5 * math.cos(0.3 + 2.0 * math.pi * q / m),
                                                   0.5 * math.sin(0.3 + 2.0 * math.pi * q / m)) for q in range(m)]),
                             density=1.5, friction=0.2, restitution=0.2)]
            kw = dict(angle=0.25 * k, angular_velocity=0.5)
        elif kind == 3:
            fx = [b2.Fixture(b2.Polygon(box=(0.45, 0.45)), density=1.0, friction=0.5, group_index=-3)]
            kw = dict(fixed_rotation=True)
        elif kind == 4:
            fx = [b2.Fixture(b2.Polygon(box=(0.6, 0.2, (0.0, 0.0), 0.3)), density=1.0, friction=0.3),
                  b2.Fixture(b2.Circle(0.25, (0.5, 0.2)), density=3.0, friction=0.3)]
            kw = dict(gravity_scale=0.5, linear_velocity=(0.5, -1.0))
        else:
            fx = [b2.Fixture(b2.Circle(0.3, (0.1, 0.0)), density=1.0, friction=0.9, category_bits=0x0002, mask_bits=0xFFFD)]
        w.create_dynamic_body(position=pos, fixtures=fx, **kw)
    return w


# ---- continuous-collision coverage scene (not a BASELINE config): bullets, thin walls, fast bodies --------------
def bullets(b2, world=0, config=7, **world_kw):
    """What World.solve_toi (world.py:594-821) reads: bullet bodies against static walls, against a kinematic
    platform and against ordinary dynamic bodies; fast non-bullet bodies against static geometry (the only TOI
    events of a scene without bullets); a thin spinning rod; circles, boxes, polygons and edges as TOI proxies.
    Continuous collision is ON by default here."""
    kw = dict(world_kw)
    kw.setdefault("continuous_physics", True)
    w = b2.World((0, -10), True, **kw)
    w.create_static_body(shapes=[b2.Edge((-20.0, 0.0), (20.0, 0.0)), b2.Edge((-12.0, 0.0), (-12.0, 12.0))])
    w.create_static_body(position=(8.0, 3.0), shapes=b2.Polygon(box=(0.05, 3.0)))
    w.create_kinematic_body(position=(-4.0, 6.0), linear_velocity=(1.5, 0.0), angular_velocity=0.3,
                            fixtures=b2.Fixture(b2.Polygon(box=(1.5, 0.1)), friction=0.4))
    n = 24
    jit = jitter(config, world, n + 3)[3:]
    k = 0

    def at(x, y):
        return (x + jit[k, 0], y + jit[k, 1])

    box = b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=1.0, friction=0.3)
    for i in range(6):                                            # a stack the bullets hit
        w.create_dynamic_body(position=at(2.0, 0.51 + 1.02 * i), fixtures=box)
        k += 1
    for i in range(4):                                            # bullets at the thin wall, the stack, the edge wall
        w.create_dynamic_body(position=at(-2.0 - 1.5 * i, 1.0 + 1.1 * i), bullet=True, linear_velocity=(70.0 + 25.0 * i, 2.0 * i),
                              fixtures=b2.Fixture(b2.Circle(0.12 + 0.03 * i), density=4.0, friction=0.2, restitution=0.3))
        k += 1
    for i in range(3):
        w.create_dynamic_body(position=at(-3.0 + 0.8 * i, 8.0 + 0.9 * i), bullet=True, linear_velocity=(-90.0, -20.0 - 10.0 * i),
                              angular_velocity=5.0, fixtures=b2.Fixture(b2.Polygon(box=(0.2, 0.1)), density=2.0, friction=0.5))
        k += 1
    for i in range(5):                                            # fast ordinary bodies: TOI against static geometry only
        m = 3 + i
        shape = b2.Polygon(vertices=[(0.4 * math.cos(0.2 + 2.0 * math.pi * q / m), 0.4 * math.sin(0.2 + 2.0 * math.pi * q / m))
                                     for q in range(m)])
        w.create_dynamic_body(position=at(-10.0 + 1.7 * i, 9.0 + 0.7 * i), linear_velocity=(3.0 - i, -45.0 - 8.0 * i), angle=0.3 * i,
                              fixtures=b2.Fixture(shape, density=1.0, friction=0.3, restitution=0.1 * i))
        k += 1
    for i in range(3):                                            # thin rods, spinning
        w.create_dynamic_body(position=at(11.0 + 2.2 * i, 5.0 + 2.0 * i), angular_velocity=25.0 - 20.0 * i, linear_velocity=(-6.0, -30.0),
                              fixtures=b2.Fixture(b2.Polygon(box=(1.2, 0.04)), density=1.0, friction=0.3))
        k += 1
    for i in range(3):                                            # bullets dropped on the kinematic platform
        w.create_dynamic_body(position=at(-5.0 + 1.2 * i, 9.0 + 1.5 * i), bullet=True, linear_velocity=(0.0, -60.0),
                              fixtures=b2.Fixture(b2.Circle(0.2), density=1.0, friction=0.3, restitution=0.5))
        k += 1
    assert k == n
    return w


# ---- sensor coverage scene (not a BASELINE config): trigger zones and proximity sensors -------------------------------------
def sensors(b2, world=0, config=8, **world_kw):
    """Sensor fixtures (fixture.py:163-168, contact.py:232-239): a static trigger zone, a sensor ring around a dynamic body (a
    second fixture, with density), a sensor on a kinematic sweeper, and ordinary bodies that fall through / pile up around them.
    Sensor contacts have no manifold, wake nobody, join no island and are never TOI candidates; their touching flag is GJK overlap."""
    w = b2.World((0, -10), True, **_kw(world_kw))
    w.create_static_body(shapes=[b2.Edge((-20.0, 0.0), (20.0, 0.0)), b2.Edge((-14.0, 0.0), (-14.0, 10.0)), b2.Edge((14.0, 0.0), (14.0, 10.0))])
    zone = w.create_static_body(position=(0.0, 3.0))
    zone.create_fixture(b2.Polygon(box=(4.0, 1.0)), sensor=True)
    zone.create_fixture(b2.Circle(1.5, (-8.0, 1.0)), sensor=True)
    # (2.3 m/s, not 2: at 2 m/s the sweeper reaches its fat box's edge -- 0.1 m margin + two steps of displacement -- after EXACTLY
    # five steps, and rounding decides the containment test)
    sweeper = w.create_kinematic_body(position=(-10.0, 1.2), linear_velocity=(2.3, 0.0))
    sweeper.create_fixture(b2.Polygon(box=(0.4, 0.4)), friction=0.5)
    sweeper.create_fixture(b2.Circle(1.6), sensor=True)
    n = 30
    jit = jitter(config, world, n + 3)[3:]
    for k in range(n):
        i, j = divmod(k, 6)
        pos = (-7.0 + 2.8 * j + jit[k, 0], 5.0 + 1.5 * i + jit[k, 1])
        if k % 5 == 0:
            body = w.create_dynamic_body(position=pos, angular_velocity=0.7)
            body.create_fixture(b2.Polygon(box=(0.4, 0.3)), density=1.0, friction=0.4)
            body.create_fixture(b2.Circle(1.1), density=0.05, sensor=True)         # proximity ring: it has mass, no response
        elif k % 5 == 1:
            w.create_dynamic_body(position=pos, fixtures=b2.Fixture(b2.Circle(0.35), density=1.0, friction=0.3, restitution=0.4))
        elif k % 5 == 2:
            m = 3 + (k // 5) % 4
            shape = b2.Polygon(vertices=[(0.45 * math.cos(0.1 + 2.0 * math.pi * q / m), 0.45 * math.sin(0.1 + 2.0 * math.pi * q / m))
                                         for q in range(m)])
            w.create_dynamic_body(position=pos, angle=0.2 * k, fixtures=b2.Fixture(shape, density=1.5, friction=0.3))
        elif k % 5 == 3:
            w.create_dynamic_body(position=pos, bullet=(k % 2 == 1), linear_velocity=(0.0, -25.0),
                                  fixtures=b2.Fixture(b2.Polygon(box=(0.3, 0.3)), density=2.0, friction=0.3))
        else:
            w.create_dynamic_body(position=pos, fixtures=b2.Fixture(b2.Polygon(box=(0.5, 0.25)), density=1.0, friction=0.6))
    return w


# ---- joint coverage scene (not a BASELINE config): revolute, distance and weld joints in the island solver -----
def joints(b2, world=0, config=9, **world_kw):
    """What the joint hooks of Island.solve (island.py:107-113, 145-148) and of the island walk (world.py:548-572) read: a plank
    bridge of revolute joints between two ground anchors (testbed/bridge.py) with bodies dropped on it, a hanging chain, a
    motorised arm with angle limits, a door with equal limits, a flail on a rigid distance joint, a soft (frequency > 0)
    distance spring, a welded cantilever, a joint with collide_connected=True, prismatic joints (motorised lift with limits, slider
    between two dynamic bodies, equal limits, lower stop), rope joints (taut and slack), a friction joint, a car on wheel joints, a pulley, a mouse joint and a joint between two dynamic bodies that start asleep."""
    w = b2.World((0, -10), True, **_kw(world_kw))
    ground = w.create_static_body(shapes=b2.Edge((-40.0, 0.0), (40.0, 0.0)))
    n_free = 12
    jit = jitter(config, world, 96)
    k = [1]

    def at(x, y):
        j = jit[k[0] % 96]
        k[0] += 1
        return (x + j[0], y + j[1])

    plank = b2.Fixture(b2.Polygon(box=(0.5, 0.125)), density=20.0, friction=0.2)
    # bridge: 10 planks from (-15, 5) to (-5, 5)
    prev = ground
    for i in range(10):
        body = w.create_dynamic_body(position=(-14.5 + 1.0 * i, 5.0), fixtures=plank)
        w.create_revolute_joint(prev, body, anchor=(-15.0 + 1.0 * i, 5.0))
        prev = body
    w.create_revolute_joint(prev, ground, anchor=(-5.0, 5.0))
    # hanging chain of 8 links from (0, 12)
    link = b2.Fixture(b2.Polygon(box=(0.6, 0.125)), density=20.0, friction=0.2)
    prev = ground
    for i in range(8):
        body = w.create_dynamic_body(position=(0.5 + 1.0 * i, 12.0), fixtures=link)
        w.create_revolute_joint(prev, body, anchor=(1.0 * i, 12.0), collide_connected=False)
        prev = body
    # motorised arm with limits; door with equal limits (a revolute that cannot turn)
    arm = w.create_dynamic_body(position=(8.0, 4.0), fixtures=b2.Fixture(b2.Polygon(box=(1.5, 0.1)), density=2.0, friction=0.3))
    w.create_revolute_joint(ground, arm, anchor=(6.5, 4.0), lower_angle=-0.4, upper_angle=0.9, limit_enabled=True,
                            motor_enabled=True, motor_speed=1.5, max_motor_torque=400.0)
    arm2 = w.create_dynamic_body(position=(-32.0, 5.0), fixtures=b2.Fixture(b2.Polygon(box=(1.5, 0.1)), density=2.0, friction=0.3))
    w.create_revolute_joint(ground, arm2, anchor=(-33.5, 5.0), lower_angle=-0.5, upper_angle=0.3, limit_enabled=True,
                            motor_enabled=True, motor_speed=-2.0, max_motor_torque=10.0)
    door = w.create_dynamic_body(position=(12.0, 3.0), fixtures=b2.Fixture(b2.Polygon(box=(0.1, 1.0)), density=1.0, friction=0.3))
    w.create_revolute_joint(ground, door, anchor=(12.0, 2.0), lower_angle=0.0, upper_angle=0.0, limit_enabled=True)
    # free-spinning motor without limit on a wheel that touches the ground (collide_connected=True keeps the contact)
    wheel = w.create_dynamic_body(position=at(16.0, 0.6), fixtures=b2.Fixture(b2.Circle(0.6), density=1.0, friction=0.9))
    axle = w.create_dynamic_body(position=(16.0, 0.6), fixtures=b2.Fixture(b2.Polygon(box=(0.2, 0.2)), density=1.0,
                                                                         category_bits=0x0004, mask_bits=0x0000))
    w.create_revolute_joint(axle, wheel, anchor=(16.0, 0.6), motor_enabled=True, motor_speed=-3.0, max_motor_torque=50.0,
                            collide_connected=True)
    # flail: rigid distance joint; spring: soft distance joint
    ball = w.create_dynamic_body(position=at(-20.0, 6.0), fixtures=b2.Fixture(b2.Circle(0.4), density=2.0, friction=0.3))
    w.create_distance_joint(ground, ball, anchor_a=(-20.0, 10.0), anchor_b=ball.position)
    swing = w.create_dynamic_body(position=at(-24.0, 7.0), linear_velocity=(3.0, 0.0),
                                  fixtures=b2.Fixture(b2.Polygon(box=(0.4, 0.4)), density=1.0, friction=0.3))
    w.create_distance_joint(ground, swing, anchor_a=(-22.0, 11.0), anchor_b=(-24.0, 7.4), frequency=2.0, damping_ratio=0.3)
    pair_a = w.create_dynamic_body(position=at(-28.0, 3.0), fixtures=b2.Fixture(b2.Circle(0.3), density=1.0, friction=0.3))
    pair_b = w.create_dynamic_body(position=at(-27.0, 4.5), fixtures=b2.Fixture(b2.Circle(0.3), density=3.0, friction=0.3))
    w.create_distance_joint(pair_a, pair_b, anchor_a=pair_a.position, anchor_b=pair_b.position, length=2.2, collide_connected=True)
    # welded cantilever of 5 beams from the ground anchor (20, 6)
    beam = b2.Fixture(b2.Polygon(box=(0.5, 0.125)), density=20.0, friction=0.3)
    prev = ground
    for i in range(5):
        body = w.create_dynamic_body(position=(20.5 + 1.0 * i, 6.0), fixtures=beam)
        w.create_weld_joint(prev, body, anchor=(20.0 + 1.0 * i, 6.0))
        prev = body
    # prismatic joints: a motorised lift with limits that a box rides on, a free slider on a tilted axis between two dynamic
    # bodies, a piston with equal limits (rigid) and a lower-limit stop
    lift = w.create_dynamic_body(position=(-36.0, 2.0), fixtures=b2.Fixture(b2.Polygon(box=(1.0, 0.2)), density=2.0, friction=0.6))
    w.create_prismatic_joint(ground, lift, anchor=(-36.0, 2.0), axis=(0.0, 1.0), lower_limit=-1.0, upper_limit=3.0,
                             limit_enabled=True, motor_enabled=True, motor_speed=1.5, max_motor_force=2000.0)
    w.create_dynamic_body(position=at(-36.0, 2.6), fixtures=b2.Fixture(b2.Polygon(box=(0.3, 0.3)), density=1.0, friction=0.6))
    rail = w.create_dynamic_body(position=at(33.0, 3.0), angle=0.3, fixtures=b2.Fixture(b2.Polygon(box=(1.5, 0.15)), density=1.0, friction=0.3))
    cart = w.create_dynamic_body(position=at(33.0, 3.6), angle=0.3, fixtures=b2.Fixture(b2.Polygon(box=(0.4, 0.3)), density=2.0, friction=0.3))
    w.create_prismatic_joint(rail, cart, anchor=(33.0, 3.3), axis=(math.cos(0.3), math.sin(0.3)), lower_limit=-1.0, upper_limit=1.0,
                             limit_enabled=True)
    piston = w.create_dynamic_body(position=(37.0, 4.0), fixtures=b2.Fixture(b2.Circle(0.4), density=1.0, friction=0.3))
    w.create_prismatic_joint(ground, piston, anchor=(37.0, 4.0), axis=(1.0, 0.0), lower_limit=0.0, upper_limit=0.0, limit_enabled=True)
    drop = w.create_dynamic_body(position=(39.0, 6.0), fixtures=b2.Fixture(b2.Polygon(box=(0.3, 0.3)), density=1.0, friction=0.3))
    w.create_prismatic_joint(ground, drop, anchor=(39.0, 6.0), axis=(0.0, 1.0), lower_limit=-2.0, upper_limit=0.5, limit_enabled=True,
                             motor_enabled=True, motor_speed=0.0, max_motor_force=1.0)
    # rope joints (a tethered ball that is thrown, a slack tether), a friction joint (top-down damping of a sliding puck) and a
    # car on two wheel joints (spring suspension, one motorised wheel) that drives along the ground
    tball = w.create_dynamic_body(position=at(-30.0, 9.0), linear_velocity=(6.0, 4.0), fixtures=b2.Fixture(b2.Circle(0.3), density=1.0, friction=0.3))
    w.create_rope_joint(ground, tball, local_anchor_a=(-30.0, 12.0), local_anchor_b=(0.0, 0.0), max_length=3.5)
    slack = w.create_dynamic_body(position=at(-26.0, 11.0), fixtures=b2.Fixture(b2.Polygon(box=(0.2, 0.2)), density=1.0, friction=0.3))
    w.create_rope_joint(ground, slack, local_anchor_a=(-26.0, 12.0), local_anchor_b=(0.1, 0.1), max_length=6.0, collide_connected=True)
    puck = w.create_dynamic_body(position=at(-18.0, 0.3), linear_velocity=(4.0, 0.0), angular_velocity=3.0,
                                 fixtures=b2.Fixture(b2.Polygon(box=(0.4, 0.25)), density=1.0, friction=0.1))
    w.create_friction_joint(ground, puck, local_anchor_a=(-18.0, 0.3), local_anchor_b=(0.0, 0.0), max_force=6.0, max_torque=2.0,
                            collide_connected=True)
    chassis = w.create_dynamic_body(position=(-2.0, 1.0), fixtures=b2.Fixture(b2.Polygon(box=(1.2, 0.25)), density=1.0, friction=0.3))
    for side, motor in ((-1.0, True), (1.0, False)):
        tyre = w.create_dynamic_body(position=at(-2.0 + 0.9 * side, 0.45), fixtures=b2.Fixture(b2.Circle(0.4), density=1.0, friction=0.9))
        w.create_wheel_joint(chassis, tyre, anchor=tyre.position, axis=(0.0, 1.0), frequency=4.0, damping_ratio=0.7,
                             motor_enabled=motor, motor_speed=-6.0, max_motor_torque=20.0)
    idler = w.create_dynamic_body(position=at(3.5, 0.5), fixtures=b2.Fixture(b2.Circle(0.4), density=1.0, friction=0.5))
    post = w.create_dynamic_body(position=(3.5, 1.5), fixtures=b2.Fixture(b2.Polygon(box=(0.2, 0.5)), density=1.0, friction=0.3))
    w.create_wheel_joint(post, idler, anchor=idler.position, axis=(0.0, 1.0), frequency=0.0, damping_ratio=0.0)   # no spring: a rigid line
    # pulley: two unequal boxes over two ground anchors, ratio 1.5
    pa_ = w.create_dynamic_body(position=at(-10.0, 14.0), fixtures=b2.Fixture(b2.Polygon(box=(0.4, 0.4)), density=2.0, friction=0.3))
    pb_ = w.create_dynamic_body(position=at(-7.0, 14.0), fixtures=b2.Fixture(b2.Polygon(box=(0.4, 0.4)), density=1.0, friction=0.3))
    w.create_pulley_joint(pa_, pb_, ground_anchor_a=(-10.0, 18.0), ground_anchor_b=(-7.0, 18.0), anchor_a=(-10.0, 14.4),
                          anchor_b=(-7.0, 14.4), ratio=1.5)
    # mouse joint: a box dragged to a target above the bridge by a soft one-body constraint
    drag = w.create_dynamic_body(position=at(-12.0, 7.0), fixtures=b2.Fixture(b2.Polygon(box=(0.3, 0.3)), density=1.0, friction=0.4))
    w.create_mouse_joint(drag, target=(-9.0, 9.5), max_force=200.0, frequency=3.0, damping_ratio=0.5)
    # two sleeping bodies joined by a revolute joint, woken when something lands on them
    sa = w.create_dynamic_body(position=(28.0, 0.26), awake=False, fixtures=b2.Fixture(b2.Polygon(box=(0.5, 0.25)), density=1.0, friction=0.4))
    sb = w.create_dynamic_body(position=(29.0, 0.26), awake=False, fixtures=b2.Fixture(b2.Polygon(box=(0.5, 0.25)), density=1.0, friction=0.4))
    w.create_revolute_joint(sa, sb, anchor=(28.5, 0.26))
    # loose bodies dropped on the bridge, the cantilever, the arm and the sleepers
    for i in range(n_free):
        x = (-13.0 + 1.6 * i) if i < 6 else (21.0 + 1.4 * (i - 6))
        if i % 3 == 0:
            shape = b2.Circle(0.35)
        elif i % 3 == 1:
            shape = b2.Polygon(box=(0.35, 0.35))
        else:
            shape = b2.Polygon(vertices=[(0.4 * math.cos(0.4 + 2.0 * math.pi * q / 5), 0.4 * math.sin(0.4 + 2.0 * math.pi * q / 5))
                                         for q in range(5)])
        w.create_dynamic_body(position=at(x, 9.0 + 0.8 * (i % 4)), fixtures=b2.Fixture(shape, density=1.0, friction=0.4))
    w.create_dynamic_body(position=at(8.5, 8.0), fixtures=b2.Fixture(b2.Polygon(box=(0.3, 0.3)), density=1.0, friction=0.5))
    w.create_dynamic_body(position=at(28.6, 5.0), fixtures=b2.Fixture(b2.Polygon(box=(0.3, 0.3)), density=2.0, friction=0.5))
    return w


# ---- joint throughput scene: a plank bridge of revolute joints under a row of boxes (testbed/bridge.py) --------------------------
def bridge(b2, world=0, planks=30, config=10, **world_kw):
    w = b2.World((0, -10), True, **_kw(world_kw))
    ground = w.create_static_body(shapes=b2.Edge((-40.0, 0.0), (40.0, 0.0)))
    plank = b2.Fixture(b2.Polygon(box=(0.5, 0.125)), density=20.0, friction=0.2)
    jit = jitter(config, world, planks + 1 + 8)
    prev = ground
    for i in range(planks):
        body = w.create_dynamic_body(position=(-14.5 + 1.0 * i, 5.0), fixtures=plank)
        w.create_revolute_joint(prev, body, anchor=(-15.0 + 1.0 * i, 5.0))
        prev = body
    w.create_revolute_joint(prev, ground, anchor=(-15.0 + 1.0 * planks, 5.0))
    for i in range(8):
        shape = b2.Circle(0.5) if i % 2 else b2.Polygon(box=(0.5, 0.5))
        j = jit[planks + 1 + i]
        w.create_dynamic_body(position=(-8.0 + 2.0 * i + j[0], 6.0 + 1.2 * (i % 3) + j[1]), fixtures=b2.Fixture(shape, density=1.0, friction=0.4))
    return w


# ---- Loop shape coverage scene (not a BASELINE config): closed chains of edges with ghost vertices ---------------------
def loops(b2, world=0, config=11, **world_kw):
    """Loop shapes (shapes.py:780-880): a static terrain loop the bodies live INSIDE (a bowl with a bump: inside collision, convex
    and concave corners), a loop on a kinematic body that turns (a tumbler: the proxies of its eight child edges move every
    step), a small loop fixture on a dynamic body next to a polygon fixture with density (loops have no mass), circles and polygons
    dropped into all of it."""
    w = b2.World((0, -10), True, **_kw(world_kw))
    terrain = [(-16.0, 0.0), (-6.0, 0.0), (-3.0, 1.5), (0.0, 0.0), (16.0, 0.0), (18.0, 8.0), (16.0, 16.0), (-16.0, 16.0), (-18.0, 8.0)]
    w.create_static_body(shapes=b2.Loop(terrain), shape_fixture=b2.Fixture(b2.Circle(1.0), friction=0.5))
    tumbler = w.create_kinematic_body(position=(8.0, 6.0), angular_velocity=0.6)
    tumbler.create_loop_fixture([(3.0 * math.cos(2.0 * math.pi * q / 8), 3.0 * math.sin(2.0 * math.pi * q / 8)) for q in range(8)], friction=0.6)
    # (an initial velocity: a body that falls from rest has moved EXACTLY the 0.1 m fat-box margin after eight steps at 60 Hz --
    # g dt^2 n (n + 1) / 2 -- and rounding decides the containment test of DynamicTree.move_proxy)
    cart = w.create_dynamic_body(position=(-10.0, 3.0), linear_velocity=(0.4, -0.3))
    cart.create_polygon_fixture(box=(1.0, 0.2), density=2.0, friction=0.4)
    cart.create_loop_fixture([(-1.0, 0.2), (1.0, 0.2), (1.0, 1.0), (-1.0, 1.0)], friction=0.4)       # an open-topped tray would be edges; this is a frame
    n = 28
    jit = jitter(config, world, n + 3)[3:]
    for k in range(n):
        i, j = divmod(k, 7)
        inside = k % 4 == 0
        pos = ((8.0 + 0.9 * (j - 3), 6.0 + 0.8 * (i - 1.5)) if inside else (-13.0 + 3.2 * j, 5.0 + 1.7 * i))
        pos = (pos[0] + jit[k, 0], pos[1] + jit[k, 1])
        if k % 3 == 0:
            shape = b2.Circle(0.3)
        elif k % 3 == 1:
            shape = b2.Polygon(box=(0.3, 0.25))
        else:
            m = 3 + (k // 3) % 4
            shape = b2.Polygon(vertices=[(0.35 * math.cos(0.2 + 2.0 * math.pi * q / m), 0.35 * math.sin(0.2 + 2.0 * math.pi * q / m)) for q in range(m)])
        w.create_dynamic_body(position=pos, angle=0.1 * k, linear_velocity=(0.0, -0.2 - 0.01 * k),      # (not from rest: see the cart)
                              fixtures=b2.Fixture(shape, density=1.0, friction=0.4, restitution=0.1 * (k % 3)))
    w.create_dynamic_body(position=(-4.0, 12.0), bullet=True, linear_velocity=(15.0, -60.0), fixtures=b2.Fixture(b2.Circle(0.15), density=4.0, friction=0.2))
    return w


# ---- controller coverage scene (not a BASELINE config): buoyancy and pairwise gravity as per-step host forces -------
def controllers(b2, world=0, config=12, **world_kw):
    """controllers.py: a pool (half plane y < 4) with buoyancy, linear and angular drag into which boxes, circles, a triangle and a
    two-fixture body are dropped, one of them asleep; three bodies without world gravity that attract each other through a
    GravityController (1 / r^2) next to it."""
    w = b2.World((0, -10), True, **_kw(world_kw))
    w.create_static_body(shapes=[b2.Edge((-20.0, 0.0), (20.0, 0.0)), b2.Edge((-12.0, 0.0), (-12.0, 10.0)), b2.Edge((12.0, 0.0), (12.0, 10.0))])
    jit = jitter(config, world, 16)
    pool = []
    for k in range(9):
        pos = (-9.0 + 2.2 * k + jit[k, 0], 5.0 + 0.9 * (k % 3) + jit[k, 1])
        if k % 4 == 0:
            fx = [b2.Fixture(b2.Circle(0.4), density=0.6, friction=0.3)]
        elif k % 4 == 1:
            fx = [b2.Fixture(b2.Polygon(box=(0.5, 0.3)), density=0.8, friction=0.3)]
        elif k % 4 == 2:
            fx = [b2.Fixture(b2.Polygon(vertices=[(-0.5, -0.3), (0.5, -0.3), (0.0, 0.6)]), density=1.5, friction=0.3)]
        else:
            fx = [b2.Fixture(b2.Polygon(box=(0.6, 0.15)), density=0.5, friction=0.3), b2.Fixture(b2.Circle(0.25, (0.5, 0.2)), density=2.0, friction=0.3)]
        pool.append(w.create_dynamic_body(position=pos, angle=0.3 * k, linear_velocity=(0.0, -0.5 - 0.1 * k), fixtures=fx))
    pool.append(w.create_dynamic_body(position=(10.5, 0.31), awake=False, fixtures=b2.Fixture(b2.Polygon(box=(0.4, 0.3)), density=0.7, friction=0.3)))
    w.create_buoyancy_controller(bodies=pool, normal=(0, 1), offset=4.0, density=2.0, velocity=(0.5, 0.0), linear_drag=2.0, angular_drag=1.0)
    stars = []
    for k, (x, y, vx, vy, r) in enumerate(((-30.0, 12.0, 0.0, 1.0, 0.6), (-26.0, 12.0, 0.0, -1.5, 0.4), (-28.0, 15.0, 1.2, 0.0, 0.3))):
        stars.append(w.create_dynamic_body(position=(x + jit[10 + k, 0], y + jit[10 + k, 1]), linear_velocity=(vx, vy), gravity_scale=0.0,
                                           fixtures=b2.Fixture(b2.Circle(r), density=2.0, friction=0.3, restitution=0.5)))
    w.create_gravity_controller(bodies=stars, G=8.0, inv_sqr=True)
    return w


SCENES = {"controllers": controllers, "loops": loops, "joints": joints, "bridge": bridge, "features": features, "pyramid": pyramid, "hello": hello, "vertical_stack": vertical_stack, "mixed": mixed,
          "deep_stacks": deep_stacks, "bullets": bullets, "sensors": sensors}
