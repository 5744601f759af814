"""Host mutation between steps: destroy_body / create_body / create_fixture / destroy_fixture on a world that is
already stepping.  tests/golden_queries/mutation.npz is the LIVE reference's trajectory of the scenario in
oracle/gen_mutation_golden.py (`script`); the same script runs here through pypybox2d_b200.World on the fp64 kernels."""
import os

import numpy as np
import pytest

import gen_mutation_golden as gm
from pypybox2d_b200 import scenes


def snapshot(world):
    st = np.array([[b.world_center.x, b.world_center.y, b._a, b.linear_velocity.x, b.linear_velocity.y, b.angular_velocity]   # (_a: the sweep angle, unwrapped like the reference's _sweep.a)
                   for b in world.bodies])
    awake = np.array([1 if b.awake else 0 for b in world.bodies], np.uint8)
    fixtures = [f for b in world.bodies for f in b.fixtures]
    fidx = {id(f): j for j, f in enumerate(fixtures)}
    con = np.array([[fidx[id(c.fixture_a)], fidx[id(c.fixture_b)], 1 if c.touching else 0, c.point_count]
                    for c in world.contact_manager.contacts], np.int32).reshape(-1, 4)
    return st, awake, con


def test_facade_mutation_guards_without_a_device():
    import pypybox2d_b200 as b2
    w = b2.World(continuous_physics=False)
    a = w.create_dynamic_body(position=(0, 1), fixtures=b2.Fixture(b2.Circle(0.5), density=1))
    f = a.create_circle_fixture(0.2, (0.5, 0), density=1.0)
    m2 = a.mass
    a.destroy_fixture(f)
    assert a.mass < m2 and f.body is None and len(a.fixtures) == 1
    with pytest.raises(ValueError):
        a.destroy_fixture(f)
    gone = []
    w.fixture_destruction_listener = gone.append            # body.py:330-333: told about every fixture of a destroyed body
    w.destroy_body(a)
    assert a not in w.bodies and a._world is None and gone == list(a.fixtures)
    with pytest.raises(ValueError):
        w.destroy_body(a)
    w.add_body(a)                                           # a destroyed body can be attached again (world.py:238-241)
    assert a in w.bodies


@pytest.mark.gpu
def test_mutation_script_matches_the_reference():
    import pypybox2d_b200 as b2
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_queries", "mutation.npz"))
    world = scenes.pyramid(b2, world=1, rows=8, real="f64")
    step = [0]
    worst = [0.0]

    def record(w):
        i = step[0]
        st, awake, con = snapshot(w)
        assert st.shape == G["s%d" % i].shape, "body count at step %d" % i
        assert np.array_equal(con, G["c%d" % i]), "contact list (fixtures, touching, point counts) at step %d" % i
        assert np.array_equal(awake, G["a%d" % i]), "awake flags at step %d" % i
        d = float(np.abs(st - G["s%d" % i]).max())
        worst[0] = max(worst[0], d)
        assert d <= 1e-9, "state at step %d: %.2e" % (i, d)
        step[0] += 1
    gm.script(b2, world, record)
    assert step[0] == int(G["n_steps"])
    print("mutation script: %d steps, worst state difference %.2e" % (step[0], worst[0]))
