"""Controllers (pypybox2d/controllers.py): host code that applies forces between the narrowphase and the island solver.

tests/golden_controllers.npz was written by oracle/gen_controllers_golden.py from the LIVE reference: the `controllers` scene (a
buoyancy pool with drag, three bodies under a GravityController) stepped 240 times, plus the forces the reference's controllers
apply to known states.  CPU: this package's controllers reproduce those forces bit for bit.  GPU: the same scene script through
pypybox2d_b200.World (fp64 kernels) follows the reference's trajectory."""
import os

import numpy as np
import pytest

import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes
from pypybox2d_b200.common import Vec2

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_controllers.npz"))


@pytest.mark.parametrize("world", [0, 1])
def test_controller_forces_match_the_reference_bitwise(world):
    w = scenes.controllers(b2, world=world)
    assert len(w.controllers) == 2 and len(w.controllers[0].bodies) == 10 and len(w.controllers[1].bodies) == 3
    checked = 0
    for probe in G["w%d_probes" % world]:
        for body, row in zip(w.bodies, probe):
            body._xf = (float(row[0]), float(row[1]), float(row[2]), float(row[3]))
            body._c, body._a = Vec2(row[4], row[5]), float(row[6])
            body._linear_velocity, body._angular_velocity = Vec2(row[7], row[8]), float(row[9])
            body._awake = bool(row[10])
            body._force, body._torque = Vec2(), 0.0
        for c in w.controllers:
            c.step(None)
        got = np.array([[b._force.x, b._force.y, b._torque] for b in w.bodies])
        assert np.array_equal(got, probe[:, 11:14]), np.abs(got - probe[:, 11:14]).max()
        checked += int((np.abs(probe[:, 11:14]).sum(axis=1) > 0).sum())
    assert checked >= 30                                  # submerged, partly submerged and dry bodies; the three stars


def test_controller_api():
    w = b2.World(continuous_physics=False)
    a = w.create_dynamic_body(position=(0, 1), fixtures=b2.Fixture(b2.Circle(0.5), density=1.0))
    c = w.create_gravity_controller(bodies=[a], G=2.0)
    assert w.controllers == [c] and c.bodies == [a] and a._controllers == [c]
    with pytest.raises(ValueError):
        c.add_body(a)
    c.remove_body(a)
    assert c.bodies == [] and a._controllers == []
    w.destroy_controller(c)
    assert w.controllers == []
    with pytest.raises(NotImplementedError):
        w.create_buoyancy_controller(use_density=True)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [0, 1])
def test_world_with_controllers_follows_the_reference(world):
    w = scenes.controllers(b2, world=world, real="f64")
    state, awake, ncontacts = G["w%d_state" % world], G["w%d_awake" % world], G["w%d_ncontacts" % world]
    worst = 0.0
    for s in range(state.shape[0]):
        w.step(1.0 / 60, 8, 3)
        got = np.array([[b._c.x, b._c.y, b._a, b._linear_velocity.x, b._linear_velocity.y, b._angular_velocity] for b in w.bodies])
        worst = max(worst, float(np.abs(got - state[s]).max()))
        assert np.abs(got - state[s]).max() < 1e-7, "step %d: %.2e" % (s, np.abs(got - state[s]).max())
        assert [int(b.awake) for b in w.bodies] == awake[s].tolist(), "awake flags, step %d" % s
        assert len(w.contact_manager.contacts) == int(ncontacts[s]), "contacts, step %d" % s
    print("controllers world %d: max |d state| vs the reference over %d steps: %.2e" % (world, state.shape[0], worst))
