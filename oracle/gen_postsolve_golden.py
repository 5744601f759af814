#!/usr/bin/env python
"""Generate tests/golden_queries/postsolve.npz from the LIVE reference (build container only): the `post_solve(contact, impulses)`
callback sequence (contact_manager.py:82-96, island.py:245-252) of an 8-row pyramid landing and settling, continuous collision
off: per call the step, the two fixture indices and the (normal, tangent) impulses of up to two manifold points."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa: E402
import refpin  # noqa: E402

DT, VI, PI, STEPS = 1.0 / 60, 8, 3, 90


def run(b2, world):
    """[(step, fixture a, fixture b, n points, n0, t0, n1, t1)] in callback order."""
    fixtures = [f for b in world.bodies for f in b.fixtures]
    fidx = {id(f): j for j, f in enumerate(fixtures)}
    rows, step = [], [0]

    def post_solve(contact, impulses):
        imp = [v for pair in impulses for v in pair] + [0.0] * 4
        rows.append((step[0], fidx[id(contact.fixture_a)], fidx[id(contact.fixture_b)], len(impulses)) + tuple(imp[:4]))
    world.contact_manager.post_solve = post_solve
    for i in range(STEPS):
        step[0] = i
        world.step(DT, VI, PI)
    return np.array(rows, np.float64).reshape(-1, 8)


def main():
    from pypybox2d_b200 import scenes
    b2 = refpin.load()
    world = scenes.pyramid(b2, world=1, rows=8)
    refpin.index_fixtures(world)
    rows = run(b2, world)
    path = os.path.join(harness._ROOT, "tests", "golden_queries", "postsolve.npz")
    np.savez_compressed(path, rows=rows)
    print("%d post_solve calls in %d steps, %d with two points -> %s (%d KB)" % (
        len(rows), STEPS, int((rows[:, 3] == 2).sum()), os.path.relpath(path, harness._ROOT), os.path.getsize(path) // 1024))


if __name__ == "__main__":
    if not refpin.available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    main()
