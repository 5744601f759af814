"""The device step SOURCE, compiled for the host (tests/hostcheck), against the oracle.

This is a debugging aid for the CPU-only container: it proves the kernel source's logic (phases, ordering,
compaction, level schedule) before any GPU time is spent.  It is NOT the product path and proves nothing
about the compiled CUDA code -- the -m gpu tests do that.
"""
import numpy as np
import pytest

import harness
import hostcheck
import oracle

INT_FIELDS = ["w_count", "f_moved", "c_fix", "c_flags", "c_mtype", "b_flags"]
FLOAT_FIELDS = ["b_pos", "b_vel", "b_xf", "b_pos0", "f_fat", "f_tight", "c_mf", "c_pts", "c_imp"]
CASES = [("hello", 0, 0.1, 10, 10, 40), ("pyramid", 1, 1 / 60, 8, 3, 120), ("vertical_stack", 1, 1 / 60, 10, 10, 150),
         ("mixed", 1, 1 / 60, 8, 3, 150), ("deep_stacks", 1, 1 / 60, 8, 3, 150), ("features", 1, 1 / 60, 8, 3, 200)]


CCD = {"continuous_physics": True}
# continuous collision on (World.solve_toi, world.py:594-821): the bullets scene has it on by construction
CCD_CASES = [("bullets", 0, 1 / 60, 8, 3, 300, {}), ("bullets", 2, 1 / 30, 10, 4, 200, {}), ("hello", 0, 0.1, 10, 10, 60, CCD),
             ("features", 1, 1 / 60, 8, 3, 150, CCD), ("pyramid", 1, 1 / 60, 8, 3, 90, CCD), ("vertical_stack", 1, 1 / 60, 10, 10, 80, CCD),
             ("sensors", 0, 1 / 60, 8, 3, 250, {}), ("sensors", 2, 1 / 60, 8, 3, 250, CCD)]


# joints in the island solver: revolute (limits, motor), distance (rigid, soft), weld; with and without CCD
# Loop shapes: one proxy and one contact slot per child edge, AABBs without the radius margin
LOOP_CASES = [("loops", 0, 1 / 60, 8, 3, 300, {}), ("loops", 1, 1 / 60, 8, 3, 200, CCD)]
JOINT_CASES = [("joints", 0, 1 / 60, 8, 3, 400, {}), ("joints", 1, 1 / 60, 10, 4, 200, {}), ("bridge", 1, 1 / 60, 8, 3, 200, {}),
               ("joints", 2, 1 / 60, 8, 3, 200, CCD)]


@pytest.mark.parametrize("scene,world,dt,vi,pi,steps,kw", [c + ({},) for c in CASES] + CCD_CASES + JOINT_CASES + LOOP_CASES)
def test_host_compiled_step_matches_oracle_f64(scene, world, dt, vi, pi, steps, kw):
    mw, OA, caps = harness.facade_arrays(scene, world, **kw)
    HA = {k: v.copy() for k, v in OA.items()}
    orc = oracle.Oracle(OA, 0)
    inv = 0.0
    hostcheck.joint_world_steps(reset=True)
    for s in range(steps):
        orc.step(dt, vi, pi)
        hostcheck.step(HA, dt, inv, vi, pi)
        inv = 1.0 / dt
        assert harness.diff_state(OA, HA, fields=INT_FIELDS) == [], "step %d" % s
        assert harness.oracle_islands(OA) == harness.oracle_islands(HA), "islands, step %d" % s
        assert harness.oracle_pairs(OA) == harness.oracle_pairs(HA), "pairs, step %d" % s
        if scene in ("joints", "bridge"):
            assert harness.diff_state(OA, HA, fields=["j_def"]) == [], "joint limit states, step %d" % s
            assert harness.oracle_island_joints(OA) == harness.oracle_island_joints(HA), "island joints, step %d" % s
            assert np.array_equal(harness.used_view(OA, "j_imp"), harness.used_view(HA, "j_imp")), "joint impulses, step %d" % s
        # BITWISE: compiled for the host the device source performs the oracle's floating-point operations one for one
        # (squares the reference writes `x ** 2` go through libm's pow on both sides, b2b_math.cuh sq_r; glibc sin / cos on
        # both sides), so every float of the state is identical after every step
        for k in FLOAT_FIELDS:
            x, y = harness.used_view(OA, k), harness.used_view(HA, k)
            assert np.array_equal(x, y), "%s, step %d: max |d| %.3e" % (k, s, np.abs(x - y).max())
    # the multi-island scenes must have gone through the joint island schedule, the single-body scene must not
    joint = hostcheck.joint_world_steps()
    if scene in ("vertical_stack", "deep_stacks"):
        assert joint >= steps // 2, joint
    if scene == "hello":
        assert joint == 0, joint


def test_settled_pile_sleep_event_keeps_the_island_lists():
    """The reference's buggy sleep rule (island.py:166-179) puts ONE body of a settled pile to sleep every 31 steps and the
    next island walk wakes it again.  The step keeps its island lists over that event when the sleeper
    is not its island's seed: same outputs as the oracle at every step, and no replan after the first step."""
    from conftest import load_golden, arrays_from_golden
    g = load_golden("pyramid_w1")
    dt, vi, pi = float(g["meta"][0]), int(g["meta"][1]), int(g["meta"][2])
    A, caps = arrays_from_golden(g)
    OA, _ = arrays_from_golden(g, prefix="ck300_", base=A)
    HA = {k: v.copy() for k, v in OA.items()}
    orc = oracle.Oracle(OA, 0)
    orc.inv_dt0 = 1.0 / dt
    hostcheck.replans(reset=True)
    awake = []
    for s in range(75):
        orc.step(dt, vi, pi)
        hostcheck.step(HA, dt, 1.0 / dt, vi, pi)
        assert harness.diff_state(OA, HA, fields=INT_FIELDS) == [], "step %d" % (300 + s)
        assert harness.oracle_islands(OA) == harness.oracle_islands(HA), "islands, step %d" % (300 + s)
        for k in FLOAT_FIELDS:
            x, y = harness.used_view(OA, k), harness.used_view(HA, k)
            assert np.array_equal(x, y), "%s, step %d" % (k, 300 + s)
        awake.append(int(((harness.used_view(HA, "b_flags") & 0x4) != 0).sum()))
    assert min(awake) == 210 and max(awake) == 211, (min(awake), max(awake))       # the event happened (twice)
    assert awake.count(210) >= 2
    assert hostcheck.replans() == 1                                                  # only the first step built the lists


@pytest.mark.parametrize("scene,world,dt,vi,pi,steps", [("pyramid", 1, 1 / 60, 8, 3, 150), ("mixed", 1, 1 / 60, 8, 3, 100),
                                                         ("features", 1, 1 / 60, 8, 3, 150), ("vertical_stack", 1, 1 / 60, 10, 10, 100)])
def test_host_compiled_fp32_step_tracks_the_oracle_one_step_at_a_time(scene, world, dt, vi, pi, steps):
    """The fp32 instantiation against the oracle, teacher-forced: from the
    oracle's fp32-rounded state ONE step agrees to fp32 accuracy and gives the same integer outputs except on tie steps.
    CPU pre-check of the arithmetic the -m gpu P-step tier grades on the device."""
    from pypybox2d_b200 import schema
    REAL = [f[0] for f in schema.FIELDS if f[3] == "real"]
    mw, OA, caps = harness.facade_arrays(scene, world)
    orc = oracle.Oracle(OA, 0)
    inv, ties, worst_p, worst_v = 0.0, 0, 0.0, 0.0
    for s in range(steps):
        T = {k: (v.astype(np.float32).astype(np.float64) if k in REAL else v.copy()) for k, v in OA.items()}
        H = {k: (v.astype(np.float32) if k in REAL else v.copy()) for k, v in T.items()}
        o2 = oracle.Oracle(T, 0)
        o2.inv_dt0 = inv
        o2.step(dt, vi, pi)
        hostcheck.step(H, dt, inv, vi, pi)
        dp = float(np.abs(harness.used_view(T, "b_pos")[:, :3] - harness.used_view(H, "b_pos")[:, :3]).max())
        dv = float(np.abs(harness.used_view(T, "b_vel")[:, :3] - harness.used_view(H, "b_vel")[:, :3]).max())
        tie = bool(harness.diff_state(T, H, fields=["w_count", "c_fix", "c_flags", "c_mtype", "b_flags"])) or \
            harness.oracle_islands(T) != harness.oracle_islands(H) or int(T["w_island"][0][3]) != int(H["w_island"][0][3])
        if tie:
            ties += 1
            assert dp <= 1e-3 and dv <= 1e-2, "step %d (tie): %.2e m, %.2e m/s" % (s, dp, dv)
        else:
            worst_p, worst_v = max(worst_p, dp), max(worst_v, dv)
            assert dp <= 5e-5 and dv <= 1e-3, "step %d: %.2e m, %.2e m/s" % (s, dp, dv)
        orc.step(dt, vi, pi)
        inv = 1.0 / dt
    assert ties <= 0.10 * steps, ties
