#!/usr/bin/env python
"""Generate tests/golden_queries/setters.npz from the LIVE reference (build container only): body properties set between steps
 -- `Body.position / angle / transform` (body.py:384-433: the proxies are moved and new contacts are looked for at
once), `linear_damping / angular_damping / gravity_scale / bullet` (body.py:710-775), `allow_sleep` (785-790), `sleeping`
(823-825), `fixed_rotation` (884-887), `mass_data` (577-604) and `Fixture.friction / restitution` (fixture.py:130-144), `Fixture.mask_bits / category_bits / group_index` (fixture.py:176-225), `Fixture.sensor` (169-174), `Body.type` (body.py:742-767), `Body.active` (848-872) on a world that is already stepping.

The script is the scenario: an 8-row pyramid (pair order pinned, continuous collision off).  After every step the body state, the
awake flags, the contact list and the broadphase tree (height, quality, preorder by fixture index) are stored;
tests/test_setters.py plays the same script through pypybox2d_b200.World.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa: E402
import refpin  # noqa: E402
import gen_mutation_golden as gm  # noqa: E402
import gen_tree_golden as gt  # noqa: E402

DT, VI, PI = 1.0 / 60, 8, 3


def script(b2, world, record, reindex=lambda w: None):
    """The scenario, written against the API module `b2` (the reference or this package)."""
    def run(n):
        for _ in range(n):
            world.step(DT, VI, PI)
            record(world)
    B = world.bodies
    run(40)
    B[5].linear_damping = 0.5
    B[6].gravity_scale = -0.5                      # floats up into the box above it
    B[7].angular_damping = 0.3
    run(30)
    B[20].position = (8.0, 6.0)                    # out of the pile: the proxy leaves its fat box, its contacts end
    run(30)
    B[20].transform = ((0.2, 9.0), 0.3)            # back above the pile, rotated
    run(40)
    p = B[3].position
    B[3].position = (p.x + 0.01, p.y)              # inside the fat box: no re-insertion
    run(20)
    B[10].fixed_rotation = True
    B[11].allow_sleep = False
    B[12].bullet = True
    run(40)
    p = B[0].position
    B[0].position = (p.x, p.y - 0.05)              # the static ground drops by half a slop-and-a-bit under the pile
    run(30)
    B[15].angle = 0.5
    run(30)
    B[8].sleeping = True                           # velocities and forces zeroed (body.py:800-812)
    B[30].transform = ((-6.0, 0.6), 0.0)
    B[31].transform = ((-6.0, 1.7), 0.0)           # two bodies moved between the same two steps, onto each other
    run(40)
    md = B[25].mass_data                           # body.py:577-604: three times as heavy, the centre of mass off the box's centre
    B[25].mass_data = type(md)(mass=3.0 * md.mass, center=(0.1, -0.05), I=3.0 * md.I + 0.2)
    B[25].linear_velocity = (0.0, 3.0)
    run(40)
    f = B[22].fixtures[0]                          # fixture.py:130-144: the contacts created from now on mix the new values
    f.friction = 0.9
    f.restitution = 0.6
    B[22].transform = ((1.0, 10.0), 0.1)           # dropped onto the pile: it bounces
    run(70)
    f14 = B[14].fixtures[0]                        # fixture.py:176-225, 269-285: collides with nothing any more, falls out of the pile
    f14.mask_bits = 0x0000
    run(25)
    B[16].fixtures[0].group_index = -3             # two boxes of one negative group stop colliding with each other
    B[17].fixtures[0].group_index = -3
    run(30)
    f14.mask_bits = 0xFFFF                         # collides again, and is dropped back onto the pile
    B[14].transform = ((0.0, 12.0), 0.0)
    B[14].linear_velocity = (0.0, 0.0)
    run(60)
    B[18].fixtures[0].category_bits = 0x0002       # B[19] no longer accepts B[18]'s category
    B[19].fixtures[0].mask_bits = 0xFFFD
    run(40)
    B[23].fixtures[0].sensor = True                # fixture.py:169-174: no collision response any more: it sinks through the pile
    run(12)
    B[23].fixtures[0].sensor = False               # solid again, inside other boxes
    run(40)
    B[22].transform = ((0.5, 14.0), 0.0)           # lifted above the pile and frozen there: body.py:742-767
    B[22].type = b2.Body.STATIC
    B[20].transform = ((0.6, 16.0), 0.2)           # dropped onto the frozen box
    B[20].linear_velocity = (0.0, 0.0)
    run(60)
    B[22].type = b2.Body.KINEMATIC                 # the platform moves away sideways under the box
    B[22].linear_velocity = (1.5, 0.0)
    run(40)
    B[22].type = b2.Body.DYNAMIC                   # and is released
    run(50)
    box = b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=5.0)
    ghost = world.create_dynamic_body(position=(0.3, 15.0), active=False, fixtures=box)    # body.py:848-872
    reindex(world)
    run(15)                                        # an inactive body has no proxies and is not simulated: it hangs there
    ghost.active = True                            # proxies created; the pair pass at the end of the next step sees them
    run(12)                                        # in free fall, no contact yet
    ghost.active = False                           # proxies and (no) contacts gone: it hangs again, keeping its velocity
    run(10)
    late = world.create_dynamic_body(position=(-0.4, 13.0), angle=0.3, active=False, fixtures=box)
    reindex(world)
    run(5)
    late.active = True
    run(70)                                        # lands on the pile


def main():
    from pypybox2d_b200 import scenes
    b2 = refpin.load()
    world = scenes.pyramid(b2, world=1, rows=8)
    refpin.index_fixtures(world)
    states, awakes, contacts, obs, pre = [], [], [], [], []

    def record(w):
        st, aw, con = gm.snapshot_reference(w)
        states.append(st); awakes.append(aw); contacts.append(con)
        bp = w.contact_manager.broadphase
        fidx = {id(f): j for j, f in enumerate([f for b in w.bodies for f in b.fixtures])}
        obs.append((bp.tree_height, bp.tree_quality))
        pre.append(np.array(gt.preorder(bp._tree._root, lambda n: fidx[id(n.user_data.fixture)]), np.int32))
    script(b2, world, record, reindex=refpin.index_fixtures)
    out = {"n_bodies": np.array([len(x) for x in states], np.int32), "states": np.concatenate(states), "awake": np.concatenate(awakes),
           "n_contacts": np.array([len(c) for c in contacts], np.int32), "contacts": np.concatenate(contacts),
           "tree_obs": np.array(obs), "tree_len": np.array([len(x) for x in pre], np.int32), "tree_pre": np.concatenate(pre)}
    path = os.path.join(harness._ROOT, "tests", "golden_queries", "setters.npz")
    np.savez_compressed(path, **out)
    print("%d steps, contacts %d..%d, awake at the end %d of %d -> %s (%d KB)" % (
        len(states), out["n_contacts"].min(), out["n_contacts"].max(), int(awakes[-1].sum()), len(awakes[-1]),
        os.path.relpath(path, harness._ROOT), os.path.getsize(path) // 1024))


if __name__ == "__main__":
    if not refpin.available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    main()
