"""Host-side drop-in surface: same names, keyword meaning and error behaviour as the reference's classes."""
import math

import pytest

import pypybox2d_b200 as b2


def test_public_names():
    for name in ("World", "Body", "Fixture", "Circle", "Polygon", "Edge", "Vec2", "LockedError"):
        assert hasattr(b2, name)
    w = b2.World((0, -10), True, continuous_physics=False)
    for m in ("create_body", "create_dynamic_body", "create_static_body", "create_kinematic_body", "step",
              "clear_forces", "bodies", "gravity", "locked", "contact_manager"):
        assert hasattr(w, m)


def test_box_mass_matches_reference_formula():
    w = b2.World(continuous_physics=False)
    body = w.create_dynamic_body(position=(1, 2), fixtures=b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=5.0))
    assert body.mass == 5.0
    assert body._I == pytest.approx(5.0 / 6.0, rel=1e-15)       # m (w^2 + h^2) / 12
    assert body._invI == 1.0 / body._I
    assert body.world_center == b2.Vec2(1, 2)
    n = body.fixtures[0].shape._normals
    assert [(v.x, v.y) for v in n] == [(0.0, -1.0), (1.0, 0.0), (0.0, 1.0), (-1.0, 0.0)]


def test_circle_mass_and_static_body():
    w = b2.World(continuous_physics=False)
    g = w.create_static_body(position=(0, -10))
    g.create_polygon_fixture(box=(50, 10))
    assert g.mass == 0.0 and g._inv_mass == 0.0
    b = w.create_dynamic_body(position=(0, 4))
    b.create_circle_fixture(1.0, density=1.0)
    assert b.mass == pytest.approx(math.pi)
    assert b.inertia == pytest.approx(0.5 * math.pi)


def test_argument_validation():
    with pytest.raises(ValueError):
        b2.Polygon(vertices=[(0, 0), (1, 0)])
    with pytest.raises(ValueError):
        b2.Polygon(vertices=[(0, 0), (0, 1), (1, 1), (1, 0)])       # clockwise
    with pytest.raises(ValueError):
        b2.Polygon(vertices=[(math.cos(0.7 * k), math.sin(0.7 * k)) for k in range(9)])   # > 8 vertices (A.9 #13)
    with pytest.raises(ValueError):
        b2.Body(position=(float("nan"), 0))
    with pytest.raises(ValueError):
        b2.Body(type=7)


def test_unsupported_features_fail_loudly():
    with pytest.raises(NotImplementedError):
        b2.World(contact_filter=lambda a, b: True)
    assert b2.Fixture(b2.Circle(1.0), sensor=True).sensor            # sensors run on the device since round 2
    w = b2.World()                                # reference default: continuous_physics=True (runs on the device)
    assert w.continuous_physics
    with pytest.raises(NotImplementedError):
        b2.World(sub_stepping=True)
    # nine of the reference's ten joint types run on the device since round 2; the gear joint still raises
    a, b = w.create_dynamic_body(position=(0, 1)), w.create_dynamic_body(position=(1, 1))
    with pytest.raises(NotImplementedError):
        w.create_gear_joint(a, b)


def test_joint_constructors_mirror_the_reference():
    """joints/revolute.py:58-93, distance.py:57-93, weld.py:56-81: anchors, reference angle and rest length from the bodies'
    transforms; World.add_joint appends to body_a's list first (world.py:268-273); destroy_body takes its joints along."""
    w = b2.World(continuous_physics=False)
    ground = w.create_static_body()
    a = w.create_dynamic_body(position=(1.0, 2.0), angle=0.5)
    b = w.create_dynamic_body(position=(3.0, 2.0))
    rj = w.create_revolute_joint(a, b, anchor=(2.0, 2.0), lower_angle=-0.1, upper_angle=0.2, limit_enabled=True)
    assert rj.body_a is a and rj.body_b is b and not rj.collide_connected and rj.limit_enabled and not rj.motor_enabled
    assert rj._reference_angle == b.angle - a.angle and rj.joint_angle == 0.0
    wa = a.get_world_point(rj.local_anchor_a)
    assert abs(wa.x - 2.0) < 1e-12 and abs(wa.y - 2.0) < 1e-12
    dj = w.create_distance_joint(ground, b, anchor_a=(3.0, 6.0), anchor_b=(3.0, 2.0))
    assert dj.length == 4.0
    with pytest.raises(ValueError):
        w.create_distance_joint(ground, b, anchor_a=(3.0, 2.0), anchor_b=(3.0, 2.0))
    wj = w.create_weld_joint(a, ground, anchor=(0.0, 0.0))
    pj = w.create_prismatic_joint(ground, b, anchor=(3.0, 2.0), axis=(0.0, 1.0), lower_limit=-1.0, upper_limit=2.0, limit_enabled=True)
    assert (pj._local_x_axis.x, pj._local_x_axis.y) == (0.0, 1.0) and pj.joint_translation == 0.0 and pj.limit_enabled
    w.destroy_joint(pj)
    rope = w.create_rope_joint(ground, b, local_anchor_a=(3.0, 6.0), max_length=5.0)
    fj = w.create_friction_joint(ground, b, max_force=2.0, max_torque=1.0)
    wheel = w.create_wheel_joint(a, b, anchor=(3.0, 2.0), axis=(0.0, 1.0), motor_enabled=True, motor_speed=1.0, max_motor_torque=5.0)
    assert rope.max_length == 5.0 and fj.max_force == 2.0 and wheel.spring_frequency == 2.0 and wheel.motor_enabled
    assert len(rope._par()) <= 12 and len(wheel._par()) <= 12
    pulley = w.create_pulley_joint(a, b, ground_anchor_a=(1.0, 6.0), ground_anchor_b=(3.0, 6.0), anchor_a=(1.0, 2.0), anchor_b=(3.0, 2.0), ratio=2.0)
    assert pulley._constant == 4.0 + 2.0 * 4.0 and pulley.ratio == 2.0
    mouse = w.create_mouse_joint(b, target=(3.0, 2.5), max_force=100.0)
    assert mouse.body_a is None and mouse.body_b is b and b.joints[-1] is mouse and mouse._par()[9] == b.mass
    mouse.target = (4.0, 3.0)
    assert mouse.target == b2.Vec2(4.0, 3.0)
    for j in (rope, fj, wheel, pulley, mouse):
        w.destroy_joint(j)
    assert w.joints == [rj, dj, wj]
    assert a.joints == [rj, wj] and b.joints == [rj, dj] and ground.joints == [dj, wj]
    with pytest.raises(ValueError):
        b2.RevoluteJoint(None, b)
    rj.motor_speed = 2.0
    assert rj.motor_speed == 2.0
    w.destroy_body(a)
    assert w.joints == [dj] and b.joints == [dj] and ground.joints == [dj]
    w.destroy_joint(dj)
    assert w.joints == [] and b.joints == []
    with pytest.raises(ValueError):
        w.destroy_joint(dj)


def test_vec2_operator_conventions():
    v = b2.Vec2(3, 4)
    assert v.length == 5.0 and v.dot((1, 1)) == 7.0
    assert v.cross(1.0) == b2.Vec2(4, -3)            # vector x scalar
    assert b2.scalar_cross(1.0, v) == b2.Vec2(-4, 3)  # scalar x vector
    y = v
    y += (1, 0)                                       # rebinding, not in place
    assert v == b2.Vec2(3, 4) and y == b2.Vec2(4, 4)


def test_default_contact_filter_rule():
    class F(object):
        def __init__(self, c, m, g):
            self._category_bits, self._mask_bits, self._group_index = c, m, g
    f = b2.World.default_contact_filter
    assert f(F(1, 0xffff, 0), F(1, 0xffff, 0))
    assert not f(F(1, 0xfffe, 0), F(1, 0xffff, 0))
    assert f(F(1, 0, 3), F(1, 0, 3)) and not f(F(1, 0xffff, -3), F(1, 0xffff, -3))


def test_create_body_copies_an_existing_body_and_inactive_bodies_pack_without_proxies():
    """World.create_body(body=...) copies the body and its fixtures (world.py:208-224, body.py:118-140); a body created with
    active=False keeps its fixtures but has no proxies (body.py:288-297): its rows can never be paired."""
    import numpy as np
    import pypybox2d_b200 as b2
    from pypybox2d_b200 import schema
    w = b2.World((0, -10), True, continuous_physics=False)
    proto = b2.Body(position=(1.0, 2.0), angle=0.25, linear_velocity=(0.5, -1.0), type=b2.Body.DYNAMIC, linear_damping=0.1,
                    fixtures=[b2.Fixture(b2.Polygon(box=(0.5, 0.25)), density=2.0, friction=0.4),
                              b2.Fixture(b2.Circle(0.2, (0.3, 0.0)), density=1.0)], user_data="proto")
    a = w.create_body(proto)
    b = w.create_dynamic_body(body=proto)
    assert a is not proto and b is not proto and a is not b and len(w.bodies) == 2
    for c in (a, b):
        assert c.type == b2.Body.DYNAMIC and c.user_data == "proto" and len(c.fixtures) == 2
        assert c.fixtures[0] is not proto.fixtures[0] and c.fixtures[0].body is c
        assert (c.position.x, c.position.y, c.angle) == (1.0, 2.0, 0.25)
        assert c.mass == proto.mass and c.inertia == proto.inertia and c.linear_damping == 0.1
    ghost = w.create_dynamic_body(position=(1.0, 2.0), active=False, fixtures=b2.Fixture(b2.Circle(0.5), density=1.0))
    assert not ghost.active
    caps = schema.Caps(4, 8, 32)
    table = schema.ShapeTable()
    A = schema.alloc(caps, 1, 8)
    schema.pack_world(w, A, 0, table)
    assert not (A["b_flags"][0, 2] & schema.BODY_ACTIVE) and (A["b_flags"][0, 0] & schema.BODY_ACTIVE)
    assert A["f_moved"][0, :5].tolist() == [1, 1, 1, 1, 0]
    assert np.all(A["f_fat"][0, 4, :2] > 1e29) and np.all(A["f_fat"][0, 4, 2:] < -1e29)
    ghost.active = True                        # before the first step the flag is host state
    schema.pack_world(w, A, 0, schema.ShapeTable())
    assert A["f_moved"][0, 4] == 1 and abs(A["f_fat"][0, 4, 0] - 0.4) < 1e-12


def test_body_frame_helpers_match_the_reference_values():
    """Body.get_world_point / get_world_vector / get_local_point / get_local_vector / velocity-of-point (body.py:662-708):
    the expected numbers are the live reference's for the same body (printed by the build container, 12 digits)."""
    w = b2.World((0, -10), True, continuous_physics=False)
    b = w.create_dynamic_body(position=(1.5, 2.0), angle=0.7, linear_velocity=(0.3, -0.2), angular_velocity=1.1,
                              fixtures=b2.Fixture(b2.Polygon(box=(0.5, 0.25)), density=2.0))
    want = [(1.47176558129, 2.499202181085), (-0.02823441871, 0.499202181085), (0.382421093642, -0.322108843619),
            (2.818119749044, 0.241249000094), (-0.8, 0.35), (0.060870773201, -0.257595250591)]
    got = [b.get_world_point((0.3, 0.4)), b.get_world_vector((0.3, 0.4)), b.get_local_point((2, 2)), b.get_local_vector((2, 2)),
           b.get_linear_velocity_from_world_point((2, 3)), b.get_linear_velocity_from_local_point((0.1, 0.2))]
    for g, e in zip(got, want):
        assert abs(g.x - e[0]) < 1e-11 and abs(g.y - e[1]) < 1e-11
    b.apply_force_to_center((1.0, 2.0))
    assert b._force == b2.Vec2(1.0, 2.0) and b._torque == 0.0
    assert b.contact_list == []
    (lx, ly), (ux, uy) = b.fixtures[0].get_aabb()
    assert lx < 1.5 < ux and ly < 2.0 < uy


def test_loop_shape_packs_one_proxy_per_child_edge():
    """shapes.py:780-880: a Loop fixture has one broadphase proxy per child edge (fixture.py:250-261), each colliding as an Edge with
    its neighbours as ghost vertices; zero mass; a child's AABB has no radius margin."""
    import numpy as np
    from pypybox2d_b200 import schema
    w = b2.World(continuous_physics=False)
    ground = w.create_static_body(shapes=b2.Loop([(-5, 0), (5, 0), (5, 5), (-5, 5)]))
    body = w.create_dynamic_body(position=(0, 2), fixtures=b2.Fixture(b2.Circle(0.5), density=1.0))
    frame = body.create_loop_fixture([(-1, -1), (1, -1), (0, 1)])
    assert ground.fixtures[0].shape.child_count == 4 and frame.shape.child_count == 3 and body.mass == pytest.approx(math.pi * 0.25)
    e = ground.fixtures[0].shape.child_edge(3)
    assert (e.vertex1, e.vertex2, e.vertex0, e.vertex3) == (b2.Vec2(-5, 5), b2.Vec2(-5, 0), b2.Vec2(5, 5), b2.Vec2(5, 0))
    with pytest.raises(ValueError):
        b2.Loop([(0, 0)])
    assert [f is ground.fixtures[0] for f in w._fixture_list()] == [True] * 4 + [False] * 4
    caps = schema.Caps(2, 8, 32)
    table = schema.ShapeTable()
    A = schema.alloc(caps, 1, 16)
    schema.pack_world(w, A, 0, table)
    assert A["w_count"][0].tolist()[:2] == [2, 8] and A["f_body"][0].tolist() == [0, 0, 0, 0, 1, 1, 1, 1]
    assert len(table) == 8 and sum((m >> 18) & 1 for m in table.meta) == 7            # seven child edges, one circle
    assert A["f_tight"][0, 0].tolist() == [-5.0, 0.0, 5.0, 0.0]                          # no radius margin (shapes.py:822-837)
    assert np.allclose(A["f_fat"][0, 0], [-5.1, -0.1, 5.1, 0.1])


def test_aabb_value_class_and_controller_exports():
    """`AABB` as callers of `World.query_aabb` build it (common.py:707-870), and the controller classes on the package."""
    import pypybox2d_b200 as b2
    a, b = b2.AABB((0, 0), (2, 1)), b2.AABB(lower_bound=(1, 0.5), upper_bound=(3, 3))
    assert a.perimeter == 6.0 and a.center.x == 1.0 and a.extents.y == 0.5 and a.valid
    assert (a + b).lower_bound.x == 0.0 and (a + b).upper_bound.y == 3.0
    assert a.test_overlap(b) and not a.contains(b) and (a + b).contains(a)
    assert not a.test_overlap(b2.AABB((2.5, 0), (3, 1)))
    a.lower_bound = (-1, -1)
    assert a.lower_bound.x == -1.0 and a.perimeter == 10.0
    assert not b2.AABB((1, 1), (0, 0)).valid
    assert issubclass(b2.GravityController, b2.Controller) and issubclass(b2.BuoyancyController, b2.Controller)


def test_transform_value_class():
    """`Body.transform` as a value with the reference's read-outs (common.py:618-706)."""
    import math
    import pypybox2d_b200 as b2
    w = b2.World(continuous_physics=False)
    a = w.create_dynamic_body(position=(1, 2), angle=0.5, fixtures=b2.Fixture(b2.Circle(0.5), density=1))
    xf = a.transform
    assert xf.position.x == 1.0 and xf.position.y == 2.0 and abs(xf.angle - 0.5) < 1e-15
    p = xf * (1.0, 0.0)
    assert abs(p.x - (1.0 + math.cos(0.5))) < 1e-15 and abs(p.y - (2.0 + math.sin(0.5))) < 1e-15
    q = xf.mul_t(p)
    assert abs(q.x - 1.0) < 1e-15 and abs(q.y) < 1e-15
    b = w.create_dynamic_body(fixtures=b2.Fixture(b2.Circle(0.5), density=1))
    b.transform = xf                                        # a transform, or (position, angle): body.py:384-395
    assert b.transform == xf and b.position.y == 2.0
    b.transform = b2.Transform(position=(4, 5), angle=0.25)
    assert b.position.x == 4.0 and abs(b.angle - 0.25) < 1e-15
