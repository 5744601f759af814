"""The oracle against the committed golden vectors (made from the LIVE reference by oracle/gen_golden.py).

Every case replays the full trajectory and demands bit-identical state digests, integer-output hashes and
integer summaries at EVERY step, and bit-identical full state at the stored checkpoints.  The scene itself is
rebuilt through this package's facade, so host-side scene construction (masses, normals, AABBs) is pinned to
the reference's initial state too.
"""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, arrays_from_golden, golden_scene_kw

import harness
import oracle
import gen_golden
from pypybox2d_b200 import schema


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_replays_reference_trajectory(case):
    g = load_golden(case)
    dt, vi, pi, steps, world = g["meta"]
    vi, pi, steps, world = int(vi), int(pi), int(steps), int(world)
    scene = str(g["scene"])
    caps = schema.Caps(*[int(x) for x in g["caps"]])
    mw, A, _ = harness.facade_arrays(scene, world, caps=caps, n_shapes=int(g["init_s_meta"].shape[0]), **golden_scene_kw(g))
    # facade-built initial state == reference-built initial state (bitwise)
    for name in gen_golden.INIT_FIELDS + [f for f in gen_golden.JOINT_INIT_FIELDS if "init_" + f in g.files]:
        ref = g["init_" + name]
        mine = harness.used_view(A, name) if not name.startswith("s_") else A[name][:ref.shape[0]]
        assert np.array_equal(np.atleast_1d(np.ascontiguousarray(mine)).view(np.uint8),
                              np.atleast_1d(np.ascontiguousarray(ref)).view(np.uint8)), name
    orc = oracle.Oracle(A, 0)
    for s in range(steps):
        assert orc.step(dt, vi, pi) == 0
        assert bytes(g["digest"][s]) == gen_golden.state_digest(A), "state digest differs at step %d" % s
        assert bytes(g["ihash"][s]) == gen_golden.int_hash(A), "integer outputs differ at step %d" % s
        assert g["ints"][s].tolist() == gen_golden.int_summary(A), "integer summary differs at step %d" % s
        for name in gen_golden.CK_FIELDS + gen_golden.JOINT_CK_FIELDS:
            key = "ck%d_%s" % (s + 1, name)
            if key in g.files:
                got = np.atleast_1d(np.ascontiguousarray(harness.used_view(A, name)))
                assert np.array_equal(got.view(np.uint8), np.atleast_1d(g[key]).view(np.uint8)), key


def test_checkpoint_restart_matches_straight_run():
    """Teacher-forcing infrastructure: restarting the oracle from a stored checkpoint reproduces the digests."""
    g = load_golden("pyramid_w1")
    dt, vi, pi, steps, world = g["meta"]
    A, caps = arrays_from_golden(g)
    A, _ = arrays_from_golden(g, prefix="ck300_", base=A)
    orc = oracle.Oracle(A, 0)
    orc.inv_dt0 = 1.0 / dt
    for s in range(300, 360):
        orc.step(dt, int(vi), int(pi))
        assert bytes(g["digest"][s]) == gen_golden.state_digest(A), s


def test_field_names_match_schema():
    names = oracle.lib().orc_field_names().decode().split(",")
    assert names == schema.FIELD_NAMES == oracle.FIELD_ORDER
