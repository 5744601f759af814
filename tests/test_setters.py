"""Body properties set between steps: `Body.position / angle / transform` (body.py:384-433), the damping, gravity
scale, bullet, allow_sleep, sleeping, fixed_rotation and mass_data setters (body.py:577-604, 710-890), `Fixture.friction /
restitution` and the collision filter (`mask_bits / category_bits / group_index`, fixture.py:130-225) on a world that is already stepping.
tests/golden_queries/setters.npz is the LIVE reference's trajectory of the scenario in oracle/gen_setters_golden.py (`script`);
the same script runs here through pypybox2d_b200.World on the fp64 kernels."""
import math
import os

import numpy as np
import pytest

import gen_setters_golden as gs
from pypybox2d_b200 import scenes
from test_mutation import snapshot
from test_tree import preorder


def test_setters_on_the_host_without_a_device():
    import pypybox2d_b200 as b2
    w = b2.World(continuous_physics=False)
    g = w.create_static_body()
    g.create_edge_fixture((-10, 0), (10, 0))
    a = w.create_dynamic_body(position=(0, 1), fixtures=b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=1))
    bp = w.contact_manager.broadphase
    leaf = w._tree_leaf[(id(a.fixtures[0]), 0)]
    box0 = w._tree._box(leaf)
    a.position = (0.05, 1.0)                                # the new tight box is inside the old fat box: dynamictree.py:164-165
    assert w._tree._box(leaf) == box0 and w._moved_proxies[(id(a.fixtures[0]), 0)][1] is False
    a.transform = ((3.0, 2.0), 0.25)
    assert w._tree._box(leaf) != box0 and w._moved_proxies[(id(a.fixtures[0]), 0)][1] is True
    assert a.position.x == 3.0 and a.position.y == 2.0 and abs(a.angle - 0.25) < 1e-15
    assert a.world_center.x == 3.0 and a._a == math.atan2(math.sin(0.25), math.cos(0.25)) and a._a0 == a._a
    a.angle = 1.0
    assert a.position.x == 3.0 and abs(a.angle - 1.0) < 1e-15 and bp.proxy_count == 2
    a.linear_damping, a.angular_damping, a.gravity_scale, a.bullet = 0.5, 0.25, 2, 1
    assert (a.linear_damping, a.angular_damping, a.gravity_scale, a.bullet) == (0.5, 0.25, 2.0, True)
    inertia = a._invI
    a.fixed_rotation = True
    assert a._invI == 0.0 and inertia > 0.0
    a.fixed_rotation = False
    assert a._invI == inertia
    a.awake = False
    a.allow_sleep = False                                   # body.py:785-790: a body that may not sleep is woken
    assert a.awake and not a.sleeping
    a.linear_velocity = (1.0, 0.0)
    a.sleeping = True
    assert not a.awake and a.linear_velocity.x == 0.0
    md = a.mass_data                                        # body.py:563-604
    assert md.mass == 1.0 and md.center.x == 0.0 and abs(md.I - 1.0 / 6.0) < 1e-15
    a.mass_data = b2.MassData(mass=2.0, center=(0.25, 0.0), I=1.0)
    assert a.mass == 2.0 and a.local_center.x == 0.25 and a._invI == 1.0 / (1.0 - 2.0 * 0.0625)
    assert abs(a.world_center.x - (3.0 + 0.25 * math.cos(1.0))) < 1e-15
    g.mass_data = b2.MassData(mass=5.0)                     # no effect on a static body
    assert g.mass == 0.0
    f = a.fixtures[0]
    f.friction, f.restitution, f.density = 0.7, 0.4, 3.0    # fixture.py:130-160
    assert (f.friction, f.restitution, f.density) == (0.7, 0.4, 3.0) and a.mass == 2.0
    a.reset_mass_data()
    assert a.mass == 3.0
    with pytest.raises(ValueError):
        f.density = -1.0
    a.linear_velocity = (1.0, 2.0)
    a.type = b2.Body.STATIC                                 # body.py:742-767
    assert a.static and a.mass == 0.0 and a.linear_velocity.x == 0.0 and a.awake
    a.type = b2.Body.DYNAMIC
    assert a.dynamic and a.mass == 3.0
    with pytest.raises(ValueError):
        a.type = 7
    f.mask_bits, f.category_bits, f.group_index = 0xFFFD, 0x0004, -2     # fixture.py:176-225 (before the first step: host fields)
    assert (f.mask_bits, f.category_bits, f.group_index) == (0xFFFD, 0x0004, -2) and not w._flagged_contacts
    f.sensor = True                                         # the body is woken (fixture.py:169-174)
    assert f.sensor and a.awake


@pytest.mark.gpu
def test_setter_script_matches_the_reference():
    import pypybox2d_b200 as b2
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_queries", "setters.npz"))
    ends = np.cumsum(G["n_contacts"])
    bends, tends = np.cumsum(G["n_bodies"]), np.cumsum(G["tree_len"])
    world = scenes.pyramid(b2, world=1, rows=8, real="f64")
    bp = world.contact_manager.broadphase
    step = [0]
    worst = [0.0]

    def record(w):
        i = step[0]
        st, awake, con = snapshot(w)
        want = G["contacts"][ends[i] - G["n_contacts"][i]:ends[i]]
        assert np.array_equal(con, want), "contact list (fixtures, touching, point counts) at step %d" % i
        b0, b1 = bends[i] - G["n_bodies"][i], bends[i]
        assert len(st) == b1 - b0, "body count at step %d" % i
        assert np.array_equal(awake, G["awake"][b0:b1]), "awake flags at step %d" % i
        d = float(np.abs(st - G["states"][b0:b1]).max())
        worst[0] = max(worst[0], d)
        # (one-ulp differences between CUDA's and glibc's sin / cos grow in the much-disturbed pile of this long script: 1e-10 after
        # 770 steps, 1e-9 after 800; contact lists, awake flags and tree topology above are compared exactly at every step)
        assert d <= 1e-7, "state at step %d: %.2e" % (i, d)
        fidx = {id(f): j for j, f in enumerate([f for b in w.bodies for f in b.fixtures])}
        name = {slot: fidx[key[0]] for key, slot in w._tree_leaf.items()}
        assert preorder(w._tree, lambda n: name[n]) == list(G["tree_pre"][tends[i] - G["tree_len"][i]:tends[i]]), \
            "tree topology at step %d" % i
        assert bp.tree_height == int(G["tree_obs"][i, 0])
        assert abs(bp.tree_quality - G["tree_obs"][i, 1]) <= 1e-9 * G["tree_obs"][i, 1]
        step[0] += 1
    gs.script(b2, world, record)
    assert step[0] == len(G["n_bodies"])
    print("setter script: %d steps, worst state difference %.2e" % (step[0], worst[0]))
