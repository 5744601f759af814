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
    with pytest.raises(NotImplementedError):
        b2.Fixture(b2.Circle(1.0), sensor=True)
    w = b2.World()                                # reference default: continuous_physics=True
    w.create_dynamic_body(position=(0, 1), fixtures=b2.Fixture(b2.Circle(0.5), density=1))
    with pytest.raises(NotImplementedError):
        w.step(1 / 60, 8, 3)
    with pytest.raises(NotImplementedError):
        w.add_joint(None)


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
