#!/usr/bin/env python
"""Generate tests/golden_queries/events.npz from the LIVE reference (build container only): the begin_contact / end_contact
callbacks (contact_manager.py:46-100, contact.py:263-266) of three scenes, as (step, kind, fixture_a, fixture_b) rows in
callback order (kind 0 = begin, 1 = end; fixture indices in world order)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa: E402
import refpin  # noqa: E402

# scene, world, scene keywords, dt, vel_iters, pos_iters, steps
CASES = [("features", 1, {}, 1.0 / 60, 8, 3, 220), ("pyramid", 1, {"rows": 8}, 1.0 / 60, 8, 3, 160),
         ("vertical_stack", 1, {}, 1.0 / 60, 10, 10, 200), ("bullets", 1, {}, 1.0 / 60, 8, 3, 200),
         ("sensors", 1, {}, 1.0 / 60, 8, 3, 300)]


def run(b2, scene, world, kw, dt, vi, pi, steps, prepare=lambda w: None):
    from pypybox2d_b200 import scenes
    w = scenes.SCENES[scene](b2, world=world, **kw)
    prepare(w)
    fixtures = [f for b in w.bodies for f in b.fixtures]
    fidx = {id(f): j for j, f in enumerate(fixtures)}
    events = []
    step = [0]
    w.contact_manager.begin_contact = lambda c: events.append((step[0], 0, fidx[id(c.fixture_a)], fidx[id(c.fixture_b)]))
    w.contact_manager.end_contact = lambda c: events.append((step[0], 1, fidx[id(c.fixture_a)], fidx[id(c.fixture_b)]))
    for s in range(steps):
        step[0] = s
        w.step(dt, vi, pi)
    return np.array(events, np.int32).reshape(-1, 4)


def main():
    b2 = refpin.load()
    out = {"n_cases": np.array(len(CASES))}
    for k, (scene, world, kw, dt, vi, pi, steps) in enumerate(CASES):
        ev = run(b2, scene, world, kw, dt, vi, pi, steps, prepare=refpin.index_fixtures)
        out["e%d" % k] = ev
        print("%s: %d steps, %d begin, %d end" % (scene, steps, int((ev[:, 1] == 0).sum()), int((ev[:, 1] == 1).sum())))
    path = os.path.join(harness._ROOT, "tests", "golden_queries", "events.npz")
    np.savez_compressed(path, **out)
    print("->", os.path.relpath(path, harness._ROOT))


if __name__ == "__main__":
    if not refpin.available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    main()
