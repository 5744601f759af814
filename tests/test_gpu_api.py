"""The drop-in surface and batch behaviour on a real GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import harness
import oracle
from pypybox2d_b200 import schema, scenes


def test_world_step_is_a_drop_in_for_the_reference_script():
    """examples/hello-bench.py's body, unchanged apart from the import (and CCD off)."""
    import pypybox2d_b200 as b2
    world = b2.World((0, -10), True, continuous_physics=False, real="f64")
    ground = world.create_static_body(position=(0, -10))
    ground.create_polygon_fixture(box=(50, 10))
    body = world.create_dynamic_body(position=(0, 4))
    body.create_circle_fixture(1.0, density=1.0)
    mw, A, caps = harness.facade_arrays("hello", 0)
    orc = oracle.Oracle(A, 0)
    for i in range(40):
        world.step(0.1, 10, 10)
        world.clear_forces()
        orc.step(0.1, 10, 10)
        assert body.position.y == pytest.approx(A["b_xf"][0, 1, 1], abs=1e-12)
        assert body.linear_velocity.y == pytest.approx(A["b_vel"][0, 1, 1], abs=1e-12)
    assert body.awake == bool(A["b_flags"][0, 1] & schema.BODY_AWAKE)
    assert len(world.contact_manager.contacts) == 1
    assert world.contact_manager.contacts[0].touching
    assert not world.locked


def test_batch_invariance_and_replicate():
    """T4: world i's trajectory does not depend on batch size, CTA placement or shard offset."""
    import torch
    import pypybox2d_b200 as b2
    w = scenes.pyramid(b2)
    big = b2.BatchedWorld.replicate(w, 300, jitter=1e-2, config=3, real="f32")
    part = b2.BatchedWorld.replicate(w, 7, jitter=1e-2, config=3, world_offset=100, real="f32", threads=32)   # a different lane-team size too
    # replicate() must build the same world as the per-world facade path
    one = scenes.pyramid(b2, world=103, config=3)
    mw = b2.BatchedWorld.from_worlds([one], real="f32", caps=big.caps)
    for name in ("b_xf", "b_pos", "f_fat", "f_tight"):
        a, b = part.download([name])[name][3], mw.download([name])[name][0]
        assert np.array_equal(a, b), name
    for bw in (big, part, mw):
        bw.step(1 / 60, 8, 3, 90)
    B, P, M = big.download(), part.download(), mw.download()
    for name in ("b_pos", "b_vel", "c_fix", "c_mtype", "c_imp", "w_count"):
        assert np.array_equal(B[name][103], P[name][3]), name
        assert np.array_equal(B[name][103], M[name][0]), name
    assert big.status() == 0
    cs = big.checksum().cpu().numpy()
    assert cs.shape == (300, 4) and np.all(cs[:, 2] > 300) and np.isfinite(cs).all()
    # the checksum rows of identical worlds are identical (the hash column is an exact integer, never NaN)
    assert np.array_equal(cs[103], part.checksum().cpu().numpy()[3])
    assert np.array_equal(cs[103], mw.checksum().cpu().numpy()[0])
    assert cs[103, 0] == np.floor(cs[103, 0]) and 0 <= cs[103, 0] < 2.0 ** 52 and len(np.unique(cs[:, 0])) == 300


def test_replicate_without_jitter_equals_the_facade_path():
    """T4 for the un-jittered batch: replicate(jitter=0) builds its proxies in the working precision like from_worlds does
    (host fp64 AABBs rounded to fp32 could differ in the last ulp and move a pair-emission tie)."""
    import pypybox2d_b200 as b2
    w = scenes.pyramid(b2)
    rep = b2.BatchedWorld.replicate(w, 5, jitter=0.0, real="f32")
    one = b2.BatchedWorld.from_worlds([scenes.pyramid(b2)], real="f32", caps=rep.caps)
    for name in ("b_xf", "b_pos", "f_fat", "f_tight", "f_moved"):
        assert np.array_equal(rep.download([name])[name][4], one.download([name])[name][0]), name
    rep.step(1 / 60, 8, 3, 60)
    one.step(1 / 60, 8, 3, 60)
    R, O = rep.download(), one.download()
    for name in ("b_pos", "b_vel", "c_fix", "c_mtype", "c_imp", "w_count", "p_pairs"):
        assert np.array_equal(R[name][4], O[name][0]), name


def test_contexts_with_different_shared_memory_sizes_coexist():
    """The dynamic shared-memory limit is a property of the kernel, not of a context: a later, smaller context must
    not lower it under an earlier context's launches (regression: `invalid argument` on the first context's step)."""
    import pypybox2d_b200 as b2
    w = scenes.pyramid(b2)
    first = b2.BatchedWorld.replicate(w, 16, jitter=1e-2, config=3, real="f32", threads=4)     # 8 worlds per warp: 20 KB
    second = b2.BatchedWorld.replicate(w, 16, jitter=1e-2, config=3, real="f32", threads=32)   # 1 world per warp: 3 KB
    first.step(1 / 60, 8, 3, 5)
    second.step(1 / 60, 8, 3, 5)
    A, B = first.download(["b_pos", "b_vel"]), second.download(["b_pos", "b_vel"])
    for name in A:
        assert np.array_equal(A[name], B[name]), name
    assert first.status() == 0 and second.status() == 0


def test_large_batch_copies_stay_bitwise_identical_from_t0():
    """Size-independent property at a batch that fills the GPU several times over (65,536 worlds = BASELINE configs[1]'s
    size): world k is a copy of distinct world k % 64, so after 120 steps from the scripted start -- free fall, landing
    transient, every kernel path (pair finding, contact creation and destruction, island rebuilds, all position
    iterations) -- every copy must equal its original bit for bit, whatever CTA, lane team, wave or chunk it ran in, and
    the fp64 build must give the oracle-checked integer outputs on every copy."""
    import torch
    import pypybox2d_b200 as b2
    w = scenes.pyramid(b2)
    seed = b2.BatchedWorld.replicate(w, 64, jitter=1e-2, config=3, real="f32")
    big = seed.tiled(65536)
    seed.step(1 / 60, 8, 3, 120)
    big.step(1 / 60, 8, 3, 120)
    assert big.status() == 0
    torch.cuda.synchronize()
    for name in ("b_pos", "b_vel", "b_flags", "c_fix", "c_flags", "c_mtype", "c_imp", "c_id", "w_count", "w_island"):
        got, ref = big.t[name], seed.t[name]
        want = ref.repeat((1024,) + (1,) * (ref.dim() - 1))
        if got.dtype.is_floating_point:
            got, want = got.view(torch.int32), want.view(torch.int32)
        assert bool((got == want).all()), name
    cs = big.checksum().cpu().numpy()
    assert np.isfinite(cs[:, 1]).all() and (cs[:, 2] > 0).all()


@pytest.mark.parametrize("scene,config,n_worlds,steps,vi,pi", [
    ("vertical_stack", 2, 65536, 60, 10, 10),      # BASELINE configs[1]: examples/verticalstack-bench.py batched to 65,536 worlds
    ("mixed", 4, 16384, 60, 8, 3),                 # configs[3] scene: circles + convex polygons, 256 bodies per world
    ("deep_stacks", 5, 32768, 60, 8, 3),           # configs[4] scene: 10 columns x 20 boxes
])
def test_other_baseline_scenes_at_scale(scene, config, n_worlds, steps, vi, pi):
    """The same size-independent property on the other BASELINE scenes: copies of 32 distinct jittered worlds, stepped
    from the scripted start inside a large batch, stay bit-identical to their originals stepped in a batch of 32."""
    import torch
    import pypybox2d_b200 as b2
    w = scenes.SCENES[scene](b2)
    seed = b2.BatchedWorld.replicate(w, 32, jitter=1e-2, config=config, real="f32")
    big = seed.tiled(n_worlds)
    seed.step(1 / 60, vi, pi, steps)
    big.step(1 / 60, vi, pi, steps)
    assert big.status() == 0 and seed.status() == 0
    torch.cuda.synchronize()
    for name in ("b_pos", "b_vel", "b_flags", "c_fix", "c_flags", "c_mtype", "c_imp", "c_id", "w_count", "w_island"):
        got, ref = big.t[name], seed.t[name]
        want = ref.repeat((n_worlds // 32,) + (1,) * (ref.dim() - 1))
        if got.dtype.is_floating_point:
            got, want = got.view(torch.int32), want.view(torch.int32)
        assert bool((got == want).all()), name
    assert int(big.t["w_count"][:, 2].min()) > 0          # contacts exist: the worlds did land


def test_step_host_matches_device_step_and_reports_overflow(monkeypatch):
    import torch
    import pypybox2d_b200 as b2
    monkeypatch.setenv("B2B_PIPE_CHUNKS", "6")           # 64 worlds in six chunks: the copy / kernel pipeline is really exercised
    w = scenes.vertical_stack(b2)
    a = b2.BatchedWorld.replicate(w, 64, jitter=1e-2, config=2, real="f32")
    b = b2.BatchedWorld.replicate(w, 64, jitter=1e-2, config=2, real="f32")
    nb = a.caps.nb
    h_in = torch.zeros((64, nb, 3), dtype=torch.float32).pin_memory()
    h_out = torch.zeros((64, nb, 6), dtype=torch.float32).pin_memory()
    for s in range(30):
        a.step_host(h_in, h_out, 1 / 60, 10, 10)
        b.step(1 / 60, 10, 10)
    B = b.download(["b_pos", "b_vel"])
    assert np.array_equal(h_out.numpy()[..., :3], B["b_pos"][..., :3])
    assert np.array_equal(h_out.numpy()[..., 3:], B["b_vel"][..., :3])
    # a force applied through the host path changes the trajectory of that body only
    h_in[5, 3, 1] = 500.0
    a.step_host(h_in, h_out, 1 / 60, 10, 10)
    b.step(1 / 60, 10, 10)
    B = b.download(["b_vel"])
    assert h_out[5, 3, 4] > B["b_vel"][5, 3, 1]
    # ... and the two paths stay bitwise equal with the force applied on both
    b.t["b_force"][5, 3, 1] = 500.0
    h_in[5, 3, 1] = 500.0
    a2 = b2.BatchedWorld.replicate(w, 64, jitter=1e-2, config=2, real="f32")
    b3 = b2.BatchedWorld.replicate(w, 64, jitter=1e-2, config=2, real="f32")
    for s in range(5):
        b3.t["b_force"][5, 3, 1] = 500.0
        a2.step_host(h_in, h_out, 1 / 60, 10, 10)
        b3.step(1 / 60, 10, 10)
    B = b3.download(["b_pos", "b_vel"])
    assert np.array_equal(h_out.numpy()[..., :3], B["b_pos"][..., :3])
    assert np.array_equal(h_out.numpy()[..., 3:], B["b_vel"][..., :3])
    # capacity overflow is reported, not silently truncated
    small = b2.BatchedWorld.from_worlds([scenes.pyramid(b2)], real="f32", caps=schema.Caps(211, 211, 300))
    small.step(1 / 60, 8, 3, 60)
    assert small.status(raise_on_overflow=False) & schema.STATUS_CONTACT_OVERFLOW
    with pytest.raises(Exception):
        small.status()


def test_world_step_raises_when_a_capacity_overflows(monkeypatch):
    """The facade never mirrors back a step that dropped contacts (the reference's lists have no capacity)."""
    import pypybox2d_b200 as b2
    from pypybox2d_b200 import batch as batch_mod
    monkeypatch.setattr(batch_mod, "world_caps", lambda worlds, **kw: schema.Caps(211, 211, 300))
    world = scenes.pyramid(b2)
    with pytest.raises(RuntimeError, match="capacity"):
        for i in range(80):
            world.step(1 / 60, 8, 3)


def test_step_host_force_wakes_a_sleeping_body():
    """Body.apply_force wakes the body (body.py:800-804); forces sent through the end-to-end entry point do the same, so a
    body that has gone to sleep moves again.  hello scene: the ball comes to rest on the box and falls asleep."""
    import torch
    import pypybox2d_b200 as b2
    bw = b2.BatchedWorld.from_worlds([scenes.SCENES["hello"](b2, world=w) for w in range(4)], real="f64")
    nb = bw.caps.nb
    bw.step(0.1, 10, 10, 200)
    D = bw.download(["b_flags", "b_pos"])
    assert not (D["b_flags"][:, 1] & schema.BODY_AWAKE).any()            # every ball sleeps
    y0 = D["b_pos"][:, 1, 1].copy()
    h_in = torch.zeros((4, nb, 3), dtype=torch.float64).pin_memory()
    h_out = torch.zeros((4, nb, 6), dtype=torch.float64).pin_memory()
    bw.step_host(h_in, h_out, 0.1, 10, 10)                                # zero force: a sleeper stays asleep
    assert not (bw.download(["b_flags"])["b_flags"][:, 1] & schema.BODY_AWAKE).any()
    h_in[2, 1, 1] = 400.0                                                 # push the ball of world 2 upwards
    bw.step_host(h_in, h_out, 0.1, 10, 10)
    D = bw.download(["b_flags", "b_pos"])
    awake = (D["b_flags"][:, 1] & schema.BODY_AWAKE) != 0
    assert awake.tolist() == [False, False, True, False]
    assert D["b_pos"][2, 1, 1] > y0[2] + 1e-3 and np.array_equal(D["b_pos"][[0, 1, 3], 1, 1], y0[[0, 1, 3]])
    assert h_out[2, 1, 1] == D["b_pos"][2, 1, 1]


def test_ragged_batch_of_different_scenes():
    """Worlds with different body / fixture / contact counts in ONE batch, each against its own oracle run."""
    import pypybox2d_b200 as b2
    names = [("hello", 0), ("features", 1), ("vertical_stack", 1), ("pyramid", 2), ("hello", 3)]
    worlds = [scenes.SCENES[n](b2, world=w) for n, w in names]
    bw = b2.BatchedWorld.from_worlds(worlds, real="f64")
    oracles = []
    for n, w in names:
        mw, A, caps = harness.facade_arrays(n, w)
        oracles.append((oracle.Oracle(A, 0), A))
    for s in range(90):
        bw.step(1 / 60, 8, 3)
        for o, A in oracles:
            o.step(1 / 60, 8, 3)
    G = bw.download()
    for i, (o, A) in enumerate(oracles):
        nb, nf, nc = [int(x) for x in A["w_count"][0][:3]]
        assert [int(x) for x in G["w_count"][i][:3]] == [nb, nf, nc], names[i]
        assert np.array_equal(G["c_fix"][i][:nc], A["c_fix"][0][:nc]), names[i]
        assert np.array_equal(G["c_mtype"][i][:nc], A["c_mtype"][0][:nc]), names[i]
        assert np.allclose(G["b_pos"][i][:nb, :3], A["b_pos"][0][:nb, :3], rtol=0, atol=1e-7), names[i]
        assert np.allclose(G["b_vel"][i][:nb, :3], A["b_vel"][0][:nb, :3], rtol=0, atol=1e-7), names[i]
    assert bw.status() == 0


def test_zero_dt_and_empty_contact_world():
    """step(0, ...) only runs the narrowphase (world.py:341-349); a world whose bodies never touch keeps no contacts."""
    import pypybox2d_b200 as b2
    w = b2.World((0, -10), True, continuous_physics=False, real="f64")
    a = w.create_dynamic_body(position=(0, 10), fixtures=b2.Fixture(b2.Circle(0.5), density=1))
    b = w.create_dynamic_body(position=(5, 10), fixtures=b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=1))
    y0 = a.position.y
    w.step(0.0, 8, 3)
    assert a.position.y == y0 and a.linear_velocity.y == 0.0
    for i in range(10):
        w.step(1 / 60, 8, 3)
    assert len(w.contact_manager.contacts) == 0
    assert a.position.y < y0 and a.position.y == pytest.approx(b.position.y, abs=1e-12)
    assert a.linear_velocity.y == pytest.approx(-10.0 * 10 / 60, abs=1e-12)


def test_solve_islands_hook_is_the_solver_stage_of_a_step():
    """T1 hook: step(0) (narrowphase alone, world.py:338-349) followed by b2b_solve_islands gives the bodies and impulses a
    full step gives -- the hook runs island construction + Island.solve and nothing else."""
    import pypybox2d_b200 as b2
    w = scenes.vertical_stack(b2)
    a = b2.BatchedWorld.replicate(w, 16, jitter=1e-2, config=2, real="f64")
    b = b2.BatchedWorld.replicate(w, 16, jitter=1e-2, config=2, real="f64")
    a.step(1 / 60, 10, 10, 40)
    b.step(1 / 60, 10, 10, 40)
    a.step(1 / 60, 10, 10)
    b.step(0.0, 10, 10)
    before = b.download(["f_fat", "w_count"])
    b.solve_islands(1 / 60, 10, 10)
    A, B = a.download(), b.download()
    for name in ("b_pos", "b_vel", "b_xf", "b_flags", "c_imp", "c_mf", "c_pts", "c_flags", "l_body", "l_contact", "w_island"):
        assert np.array_equal(A[name], B[name]), name
    # ... and nothing after it: proxies not synchronised, no new contacts
    assert np.array_equal(before["f_fat"], B["f_fat"]) and np.array_equal(before["w_count"][:, 2], B["w_count"][:, 2])
    assert a.status() == 0 and b.status() == 0


def test_inactive_body_is_never_simulated():
    """Body(active=False): no proxies, no contacts, no island -- it stays where it is while an active twin falls onto the
    ground; the world around it steps exactly like a world without it."""
    import pypybox2d_b200 as b2

    def build(with_ghost):
        w = b2.World((0, -10), True, continuous_physics=False, real="f64")
        w.create_static_body(shapes=b2.Edge((-20, 0), (20, 0)))
        w.create_dynamic_body(position=(0, 3), fixtures=b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=1))
        if with_ghost:
            w.create_dynamic_body(position=(0.2, 1.0), active=False, fixtures=b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=1))
        return w
    w1, w0 = build(True), build(False)
    for i in range(90):
        w1.step(1 / 60, 8, 3)
        w0.step(1 / 60, 8, 3)
    ghost = w1.bodies[2]
    assert (ghost.position.x, ghost.position.y) == (0.2, 1.0) and ghost.linear_velocity.y == 0.0
    assert w1.bodies[1].position.y == w0.bodies[1].position.y < 0.6           # fell THROUGH the ghost onto the ground
    assert len(w1.contact_manager.contacts) == len(w0.contact_manager.contacts) == 1


def test_default_world_with_continuous_collision_is_a_drop_in():
    """The reference's DEFAULT world has continuous collision on (world.py:65-78): an unmodified user script -- here a bullet
    fired at a thin wall and a fast box dropped on the ground -- steps through World.solve_toi on the device and lands
    where the oracle's does."""
    import pypybox2d_b200 as b2
    mw, A, caps = harness.facade_arrays("bullets", 1)
    orc = oracle.Oracle(A, 0)
    world = scenes.bullets(b2, world=1, real="f64")
    assert world.continuous_physics
    for i in range(120):
        world.step(1 / 60, 8, 3)
        orc.step(1 / 60, 8, 3)
    nb = len(world.bodies)
    got = np.array([[b.position.x, b.position.y, b.angle] for b in world.bodies])
    want_xf = A["b_xf"][0, :nb]
    assert np.abs(got[:, 0] - want_xf[:, 0]).max() < 1e-6 and np.abs(got[:, 1] - want_xf[:, 1]).max() < 1e-6
    assert [b.awake for b in world.bodies] == [bool(f & schema.BODY_AWAKE) for f in A["b_flags"][0, :nb]]
    assert len(world.contact_manager.contacts) == int(A["w_count"][0][2])
    # the same script with continuous collision switched off lets the bullets tunnel: a different trajectory
    off = scenes.bullets(b2, world=1, real="f64", continuous_physics=False)
    for i in range(120):
        off.step(1 / 60, 8, 3)
    assert max(abs(a.position.x - b.position.x) for a, b in zip(off.bodies, world.bodies)) > 1.0


def test_continuous_collision_batch_invariance_and_host_path(monkeypatch):
    """T4 with World.solve_toi in the step: world i of a large fp32 batch equals the same world stepped alone, the host-buffer
    entry point reads the state back AFTER the TOI kernel, and worlds with continuous collision off in the same batch are
    untouched by it."""
    import torch
    import pypybox2d_b200 as b2
    monkeypatch.setenv("B2B_PIPE_CHUNKS", "5")
    w = scenes.bullets(b2)
    big = b2.BatchedWorld.replicate(w, 200, jitter=1e-2, config=7, real="f32")
    part = b2.BatchedWorld.replicate(w, 3, jitter=1e-2, config=7, world_offset=150, real="f32")
    host = b2.BatchedWorld.replicate(w, 200, jitter=1e-2, config=7, real="f32")
    assert big.continuous and host.continuous
    nb = big.caps.nb
    h_out = torch.zeros((200, nb, 6), dtype=torch.float32).pin_memory()
    for s in range(90):
        big.step(1 / 60, 8, 3)
        part.step(1 / 60, 8, 3)
        host.step_host(None, h_out, 1 / 60, 8, 3)
    B, P = big.download(), part.download()
    for name in ("b_pos", "b_vel", "b_pos0", "c_fix", "c_mtype", "c_imp", "w_count", "b_flags"):
        assert np.array_equal(B[name][151], P[name][1]), name
    assert np.array_equal(h_out.numpy()[..., :3], B["b_pos"][..., :3])
    assert np.array_equal(h_out.numpy()[..., 3:], B["b_vel"][..., :3])
    assert big.status() == 0
    # every jittered world is a different trajectory (the bullets hit the stack differently)
    assert len(np.unique(big.checksum().cpu().numpy()[:, 0])) == 200
    # a mixed batch: continuous collision on in the even worlds only
    mixed = b2.BatchedWorld.replicate(w, 8, jitter=1e-2, config=7, real="f32")
    mixed.t["w_flags"][1::2] &= ~schema.WORLD_CONTINUOUS
    offw = b2.BatchedWorld.replicate(scenes.bullets(b2, continuous_physics=False), 8, jitter=1e-2, config=7, real="f32")
    for s in range(60):
        mixed.step(1 / 60, 8, 3)
        offw.step(1 / 60, 8, 3)
    M, O = mixed.download(["b_pos"]), offw.download(["b_pos"])
    assert np.array_equal(M["b_pos"][1::2, :, :3], O["b_pos"][1::2, :, :3])
    assert not np.array_equal(M["b_pos"][0::2, :, :3], O["b_pos"][0::2, :, :3])


def test_world_with_joints_is_a_drop_in():
    """a scene script that creates revolute / distance / weld joints steps on the device through the facade; body
    state, joint limit states and accumulated joint impulses follow the oracle's (fp64: to the one-ulp libm differences), and a
    joint created or destroyed on a stepping world re-packs it."""
    import pypybox2d_b200 as b2
    world = scenes.joints(b2, world=0, real="f64")
    mw, A, caps = harness.facade_arrays("joints", 0)
    orc = oracle.Oracle(A, 0)
    nb, nj = len(world.bodies), len(world.joints)
    for i in range(120):
        world.step(1 / 60, 8, 3)
        orc.step(1 / 60, 8, 3)
    pos = np.array([[b._c.x, b._c.y, b._a] for b in world.bodies])
    assert np.abs(pos - A["b_pos"][0, :nb, :3]).max() < 1e-6
    assert [j._limit_state for j in world.joints] == [(int(f) >> schema.JOINT_LIMIT_SHIFT) & 3 for f in A["j_def"][0, :nj, 3]]
    assert max(j._limit_state for j in world.joints) > 0                         # an arm sits on a limit
    imp = np.array([list(j._imp()) for j in world.joints])
    assert np.abs(imp - A["j_imp"][0, :nj]).max() < 1e-5 * max(1.0, np.abs(imp).max())
    arm = world.joints[19]
    assert arm.get_motor_torque(60.0) != 0.0 and abs(arm.joint_angle) > 0.1
    # host mutation: cut the chain, hang a new body on a new joint
    chain_joint = world.joints[14]
    world.destroy_joint(chain_joint)
    extra = world.create_dynamic_body(position=(4.0, 14.0), fixtures=b2.Fixture(b2.Circle(0.3), density=1.0))
    world.create_distance_joint(world.bodies[0], extra, anchor_a=(4.0, 16.0), anchor_b=(4.0, 14.0))
    def gap(j):
        pa, pb = j.anchor_a, j.anchor_b
        return float(np.hypot(pa.x - pb.x, pa.y - pb.y))
    assert gap(chain_joint) < 0.05 and all(gap(j) < 0.05 for j in world.joints[11:18])
    for i in range(60):
        world.step(1 / 60, 8, 3)
    assert gap(chain_joint) > 0.5                                                # the cut end of the chain falls away
    assert all(gap(j) < 0.05 for j in world.joints[11:18])                       # ... the joints that remain hold
    assert abs(extra.position.y - 14.0) < 0.05 and len(world.joints) == nj       # the new body hangs on its joint


def test_joint_batches_fp32_invariance_and_sleeping_bridge():
    """Batches with a joint capacity run the solve kernels with the joint hooks (fp32): world i's trajectory does not depend on
    batch size, shard offset or lane-team width (T4), worlds without joints in the same batch step as before, and the bridge
    keeps its joints together."""
    import pypybox2d_b200 as b2
    w = scenes.bridge(b2)
    big = b2.BatchedWorld.replicate(w, 200, jitter=1e-2, config=10, real="f32")
    part = b2.BatchedWorld.replicate(w, 5, jitter=1e-2, config=10, world_offset=60, real="f32", threads=32)
    assert big.caps.nj == 31
    for bw in (big, part):
        bw.step(1 / 60, 8, 3, 150)
    B, P = big.download(), part.download()
    for name in ("b_pos", "b_vel", "c_fix", "c_mtype", "c_imp", "w_count", "j_imp", "j_def", "l_joint"):
        assert np.array_equal(B[name][62], P[name][2]), name
    assert big.status() == 0
    # every joint holds: the anchors of consecutive planks coincide to the solver's tolerance
    xf = B["b_xf"].astype(np.float64)
    d = B["j_def"]
    par = B["j_par"].astype(np.float64)
    wi = np.arange(200)[:, None]
    ba, bb = d[:, :, 1], d[:, :, 2]
    def anchor(body, lx, ly):
        t = xf[wi, body]
        return t[..., 0] + t[..., 2] * lx - t[..., 3] * ly, t[..., 1] + t[..., 3] * lx + t[..., 2] * ly
    ax, ay = anchor(ba, par[..., 0], par[..., 1])
    bx, by = anchor(bb, par[..., 2], par[..., 3])
    assert np.hypot(ax - bx, ay - by).max() < 0.05
    # a ragged batch: a world with joints next to worlds without (they share the joint-capable kernels)
    mixed = b2.BatchedWorld.from_worlds([scenes.bridge(b2, world=1), scenes.hello(b2), scenes.pyramid(b2, rows=6)], real="f64")
    refs = [harness.facade_arrays(n, wd, **kw) for n, wd, kw in (("bridge", 1, {}), ("hello", 0, {}), ("pyramid", 0, {"rows": 6}))]
    orcs = [oracle.Oracle(r[1], 0) for r in refs]
    for s in range(100):
        mixed.step(1 / 60, 8, 3)
        for o in orcs:
            o.step(1 / 60, 8, 3)
    M = mixed.download()
    for i, r in enumerate(refs):
        n = int(r[1]["w_count"][0][0])
        assert np.abs(M["b_pos"][i, :n, :3] - r[1]["b_pos"][0, :n, :3]).max() < 1e-7, i
        assert int(M["w_count"][i][2]) == int(r[1]["w_count"][0][2])


def test_world_with_loop_shapes_is_a_drop_in():
    """Loop shapes through the facade -- bodies inside a terrain loop and a turning loop on a kinematic body follow the
    oracle; contacts and queries name the Loop fixture once per child edge, as the reference's broadphase does."""
    import pypybox2d_b200 as b2
    world = scenes.loops(b2, world=0, real="f64")
    mw, A, caps = harness.facade_arrays("loops", 0)
    orc = oracle.Oracle(A, 0)
    nb = len(world.bodies)
    for i in range(120):
        world.step(1 / 60, 8, 3)
        orc.step(1 / 60, 8, 3)
    pos = np.array([[b._c.x, b._c.y, b._a] for b in world.bodies])
    assert np.abs(pos - A["b_pos"][0, :nb, :3]).max() < 1e-6
    contacts = world.contact_manager.contacts
    assert len(contacts) == int(A["w_count"][0][2]) and sum(c.touching for c in contacts) > 10
    terrain = world.bodies[0].fixtures[0]
    assert any(c.fixture_a is terrain for c in contacts)                             # the loop is fixture A (contact.py:310-311)
    hits = list(world.query_aabb(((-20.0, -1.0), (20.0, 0.5))))
    assert hits.count(terrain) >= 4                                                  # one hit per child edge whose fat box overlaps
    rc = [h for h in world.ray_cast((0.0, 10.0), (0.0, -5.0))]
    assert any(h[0] is terrain and abs(h[1].y) < 1e-9 for h in rc)                   # the ray reaches the floor edge of the loop
