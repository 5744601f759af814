#!/usr/bin/env python
"""Generate tests/golden_queries/mutation.npz from the LIVE reference (build container only): host mutation between steps
 -- World.destroy_body, World.create_dynamic_body, Body.create_fixture and Body.destroy_fixture on a world that
is already stepping (world.py:183-251, body.py:164-286).

The script is the scenario: an 8-row pyramid (pair order pinned, continuous collision off) is stepped, a box in the middle of
the pile is destroyed, a new box and a new two-fixture body are dropped in, a fixture is removed from that body.  After every
step the body state (c, a, v, w per body, in world.bodies order), the contact list (fixture indices, touching flags, manifold
point counts) and the awake flags are stored; tests/test_mutation.py plays the same script through pypybox2d_b200.World.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa: E402
import refpin  # noqa: E402

DT, VI, PI = 1.0 / 60, 8, 3


def script(b2, world, record, reindex=lambda w: None):
    """The scenario, written against the API module `b2` (the reference or this package)."""
    def run(n):
        for _ in range(n):
            world.step(DT, VI, PI)
            record(world)
    run(70)
    world.destroy_body(world.bodies[12])                       # a box of the second row, loaded from above
    reindex(world)
    run(40)
    box = b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=5.0)
    world.create_dynamic_body(position=(0.3, 12.0), fixtures=box)
    reindex(world)
    run(25)
    two = world.create_dynamic_body(position=(-3.0, 11.0), angle=0.4)
    two.create_polygon_fixture(box=(0.6, 0.2), density=2.0, friction=0.4)
    lump = two.create_circle_fixture(0.3, (0.5, 0.1), density=3.0)
    reindex(world)
    run(50)
    two.destroy_fixture(lump)
    reindex(world)
    run(40)
    world.destroy_body(world.bodies[1])                        # a corner box of the bottom row
    reindex(world)
    run(30)


def snapshot_reference(w):
    bodies = w.bodies
    fixtures = [f for b in bodies for f in b.fixtures]
    fidx = {id(f): j for j, f in enumerate(fixtures)}
    st = np.array([[b._sweep.c.x, b._sweep.c.y, b._sweep.a, b._linear_velocity.x, b._linear_velocity.y, b._angular_velocity]
                   for b in bodies])
    awake = np.array([1 if b._awake else 0 for b in bodies], np.uint8)
    con = np.array([[fidx[id(c._fixture_a)], fidx[id(c._fixture_b)], 1 if c._touching_flag else 0, c._manifold.point_count]
                    for c in w.contact_manager.contacts], np.int32).reshape(-1, 4)
    return st, awake, con


def main():
    from pypybox2d_b200 import scenes
    b2 = refpin.load()
    world = scenes.pyramid(b2, world=1, rows=8)
    refpin.index_fixtures(world)
    states, awakes, contacts = [], [], []

    def record(w):
        st, aw, con = snapshot_reference(w)
        states.append(st); awakes.append(aw); contacts.append(con)
    script(b2, world, record, reindex=refpin.index_fixtures)
    out = {"n_steps": np.array(len(states))}
    for i, (st, aw, con) in enumerate(zip(states, awakes, contacts)):
        out["s%d" % i] = st
        out["a%d" % i] = aw
        out["c%d" % i] = con
    path = os.path.join(harness._ROOT, "tests", "golden_queries", "mutation.npz")
    np.savez_compressed(path, **out)
    print("%d steps, bodies %d -> %d, contacts at the end %d -> %s (%d KB)" % (
        len(states), len(states[0]), len(states[-1]), len(contacts[-1]), os.path.relpath(path, harness._ROOT), os.path.getsize(path) // 1024))


if __name__ == "__main__":
    if not refpin.available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    main()
