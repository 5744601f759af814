#!/usr/bin/env python
"""tests/golden_controllers.npz from the LIVE reference (build container only): the `controllers` scene (scenes.py: a buoyancy
pool and three bodies under a GravityController) stepped by the unmodified pure-Python reference with the pinned pair order; the
body state (c, a, v, w), awake flags and contact counts after every step.  tests/test_controllers.py plays the same scene script
through pypybox2d_b200.World on the GPU.  (The C oracle has no controllers: they are host code that feeds forces into the step.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa: E402
import refpin  # noqa: E402
from pypybox2d_b200 import scenes  # noqa: E402

STEPS = 240
out = {}
for world in (0, 1):
    b2 = refpin.load()
    rw = scenes.controllers(b2, world=world)
    refpin.index_fixtures(rw)
    nb = len(rw.bodies)
    state = np.zeros((STEPS, nb, 6))
    awake = np.zeros((STEPS, nb), np.uint8)
    ncontacts = np.zeros(STEPS, np.int32)
    probes = []      # (step, per body: xf(4), c(2), a, v(2), w, awake, force(2), torque) -- the controllers' forces on a known state
    for s in range(STEPS):
        if s in (0, 20, 45, 90, 150, 239):
            for c in rw.controllers:
                c.step(None)
            rows = []
            for b in rw.bodies:
                xf = b._xf
                rows.append([xf._position.x, xf._position.y, xf._rotation._col1.x, xf._rotation._col1.y, b._sweep.c.x, b._sweep.c.y,
                             b._sweep.a, b._linear_velocity.x, b._linear_velocity.y, b._angular_velocity, 1.0 if b._awake else 0.0,
                             b._force.x, b._force.y, b._torque])
            probes.append(np.array(rows))
            rw.clear_forces()
            if s == 0:
                for b in rw.bodies:
                    b._awake = bool(rows[rw.bodies.index(b)][10]) if False else b._awake
        rw.step(1.0 / 60, 8, 3)
        for i, b in enumerate(rw.bodies):
            state[s, i] = (b._sweep.c.x, b._sweep.c.y, b._sweep.a, b._linear_velocity.x, b._linear_velocity.y, b._angular_velocity)
            awake[s, i] = 1 if b._awake else 0
        ncontacts[s] = len(rw.contact_manager.contacts)
    out["w%d_state" % world], out["w%d_awake" % world], out["w%d_ncontacts" % world] = state, awake, ncontacts
    out["w%d_probes" % world] = np.array(probes)
    print("world %d: %d bodies, %d steps, final contacts %d, awake %d, y range %.2f .. %.2f" % (
        world, nb, STEPS, ncontacts[-1], int(awake[-1].sum()), state[-1, 1:, 1].min(), state[-1, 1:, 1].max()))
np.savez_compressed(os.path.join(harness._ROOT, "tests", "golden_controllers.npz"), **out)
