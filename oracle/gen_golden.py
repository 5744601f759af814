#!/usr/bin/env python
"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

For every (scene, world) case the unmodified pure-Python reference (pair order pinned, CCD off --
see oracle/refpin.py) and the C oracle are stepped side by side from the same initial state.  The
script FAILS if any used row of any state field, any island list or any emitted pair list differs
bit-for-bit at any step: that is what pins the oracle.  It then writes a small fixture:

  init_*           initial schema arrays as extracted from the reference objects (used rows)
  digest[steps]    sha256 over <6d (c.x, c.y, a, v.x, v.y, w) of every body, per step
  ints[steps, 8]   n_contacts, n_touching, n_points, n_islands, n_island_bodies, n_island_contacts,
                   n_pairs, n_awake -- the graded integer outputs in summary
  ihash[steps]     sha256 over contact fixture ids + flags + manifold type/count/ids + island lists + pairs
  ck<k>_*          full used-row state after k steps, for teacher forcing / settled starts

Usage: python oracle/gen_golden.py [case ...]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa: E402
import oracle  # noqa: E402
import refpin  # noqa: E402
from pypybox2d_b200 import schema  # noqa: E402

OUT = os.path.join(harness._ROOT, "tests", "golden")

CCD = {"continuous_physics": True}

# name: (scene, world, dt, vel_iters, pos_iters, steps, checkpoints[, scene keywords])
CASES = {
    "pyramid_w0": ("pyramid", 0, 1.0 / 60, 8, 3, 600, (1, 60, 300, 600)),
    "pyramid_w1": ("pyramid", 1, 1.0 / 60, 8, 3, 600, (1, 30, 300, 600)),
    "pyramid_w2_10_10": ("pyramid", 2, 1.0 / 60, 10, 10, 200, (200,)),
    "hello_w0": ("hello", 0, 0.1, 10, 10, 200, (20, 200)),
    "vertical_stack_w0": ("vertical_stack", 0, 1.0 / 60, 10, 10, 400, (100, 400)),
    "vertical_stack_w1": ("vertical_stack", 1, 1.0 / 60, 10, 10, 400, (100, 400)),
    "mixed_w0": ("mixed", 0, 1.0 / 60, 8, 3, 400, (100, 400)),
    "mixed_w1": ("mixed", 1, 1.0 / 60, 8, 3, 400, (100, 400)),
    "deep_stacks_w1": ("deep_stacks", 1, 1.0 / 60, 8, 3, 400, (150, 400)),
    "features_w0": ("features", 0, 1.0 / 60, 8, 3, 400, (100, 400)),
    "features_w1": ("features", 1, 1.0 / 60, 10, 4, 400, (400,)),
    # continuous collision ON (World.solve_toi, world.py:594-821): the reference's default world
    "hello_ccd_w0": ("hello", 0, 0.1, 10, 10, 200, (20, 200), CCD),
    "bullets_w0": ("bullets", 0, 1.0 / 60, 8, 3, 300, (10, 300)),
    "bullets_w1": ("bullets", 1, 1.0 / 60, 8, 3, 300, (10, 300)),
    "bullets_w2": ("bullets", 2, 1.0 / 30, 10, 4, 200, (200,)),
    "pyramid_ccd_w1": ("pyramid", 1, 1.0 / 60, 8, 3, 300, (60, 300), CCD),
    "vertical_stack_ccd_w1": ("vertical_stack", 1, 1.0 / 60, 10, 10, 200, (200,), CCD),
    "features_ccd_w1": ("features", 1, 1.0 / 60, 8, 3, 300, (300,), CCD),
    "mixed_ccd_w1": ("mixed", 1, 1.0 / 60, 8, 3, 200, (200,), CCD),
    # sensor fixtures (contact.py:232-239, distance.test_overlap): trigger zones, proximity rings, with and without CCD
    "sensors_w0": ("sensors", 0, 1.0 / 60, 8, 3, 300, (100, 300)),
    "sensors_w1": ("sensors", 1, 1.0 / 60, 8, 3, 300, (300,)),
    "sensors_ccd_w2": ("sensors", 2, 1.0 / 60, 8, 3, 300, (300,), CCD),
    # joints in the island solver: revolute (limits, motor), distance (rigid, soft), weld
    "joints_w0": ("joints", 0, 1.0 / 60, 8, 3, 400, (100, 400)),
    "joints_w1": ("joints", 1, 1.0 / 60, 10, 4, 300, (300,)),
    "joints_ccd_w2": ("joints", 2, 1.0 / 60, 8, 3, 300, (300,), CCD),
    "bridge_w1": ("bridge", 1, 1.0 / 60, 8, 3, 300, (150, 300)),
    # Loop shapes: one proxy and one contact slot per child edge
    "loops_w0": ("loops", 0, 1.0 / 60, 8, 3, 300, (100, 300)),
    "loops_ccd_w1": ("loops", 1, 1.0 / 60, 8, 3, 300, (300,), CCD),
}

CK_FIELDS = ["w_count", "w_flags", "w_island", "b_xf", "b_pos", "b_vel", "b_pos0", "b_force", "b_flags", "b_island",
             "f_fat", "f_tight", "f_moved", "c_fix", "c_flags", "c_mat", "c_mtype", "c_mf", "c_pts", "c_imp", "c_id"]
INIT_FIELDS = [f[0] for f in schema.FIELDS if f[0][0] in "wbfc" or f[0].startswith("s_")]
JOINT_INIT_FIELDS = ["w_joint", "j_def", "j_par", "j_imp", "j_adj", "b_jadj"]      # only stored by fixtures of scenes with joints
INIT_FIELDS = [f for f in INIT_FIELDS if f not in JOINT_INIT_FIELDS]
JOINT_CK_FIELDS = ["j_def", "j_imp"]


def state_digest(A, w=0):
    nb = int(A["w_count"][w][0])
    buf = np.concatenate([A["b_pos"][w][:nb, :3], A["b_vel"][w][:nb, :3]], axis=1).astype("<f8")
    h = hashlib.sha256(np.ascontiguousarray(buf).tobytes())
    nj = int(A["w_joint"][w][0])
    if nj:
        h.update(np.ascontiguousarray(A["j_imp"][w][:nj].astype("<f8")).tobytes())
    return h.digest()


def int_summary(A, w=0):
    nb, nf, nc, npair = [int(x) for x in A["w_count"][w]]
    fl = A["c_flags"][w][:nc]
    mt = A["c_mtype"][w][:nc]
    touching = (fl & schema.CONTACT_TOUCHING) != 0
    npts = int(((mt >> 8) & 0xff)[touching].sum())
    isl = A["w_island"][w]
    awake = int(((A["b_flags"][w][:nb] & schema.BODY_AWAKE) != 0).sum())
    return [nc, int(touching.sum()), npts, int(isl[0]), int(isl[1]), int(isl[2]), npair, awake]


def int_hash(A, w=0):
    nb, nf, nc, npair = [int(x) for x in A["w_count"][w]]
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(A["c_fix"][w][:nc]).tobytes())
    h.update(np.ascontiguousarray(A["c_flags"][w][:nc] & np.uint32(~schema.CONTACT_ISLAND & 0xffffffff)).tobytes())
    mt = A["c_mtype"][w][:nc]
    h.update(np.ascontiguousarray(mt).tobytes())
    cnt = (mt >> 8) & 0xff
    ids = A["c_id"][w][:nc].copy()
    ids[cnt < 1, 0] = 0          # ids beyond point_count are stale storage, not outputs
    ids[cnt < 2, 1] = 0
    h.update(np.ascontiguousarray(ids).tobytes())
    h.update(np.ascontiguousarray(A["b_flags"][w][:nb] & np.uint32(~schema.BODY_ISLAND & 0xffffffff)).tobytes())
    n_isl, nlb, nlc, _ = [int(x) for x in A["w_island"][w]]
    h.update(np.ascontiguousarray(A["l_body"][w][:nlb]).tobytes())
    h.update(np.ascontiguousarray(A["l_contact"][w][:nlc]).tobytes())
    h.update(np.ascontiguousarray(A["l_start"][w][:n_isl]).tobytes())
    h.update(np.ascontiguousarray(A["p_pairs"][w][:min(npair, A["p_pairs"].shape[1])]).tobytes())
    nj = int(A["w_joint"][w][0])
    if nj:                                                  # scenes with joints: limit states, island joint lists
        h.update(np.ascontiguousarray(A["j_def"][w][:nj] & np.int32(~schema.JOINT_ISLAND)).tobytes())
        h.update(np.ascontiguousarray(A["l_joint"][w][:int(A["w_joint"][w][1])]).tobytes())
        h.update(np.ascontiguousarray(A["l_jstart"][w][:n_isl]).tobytes())
    return h.digest()


def used(A, name, w=0):
    return np.array(harness.used_view(A, name, w), copy=True, order="C")


def generate(case):
    scene, world, dt, vi, pi, steps, cks = CASES[case][:7]
    kw = CASES[case][7] if len(CASES[case]) > 7 else {}
    rw, RA, caps, table = harness.reference_arrays(scene, world, **kw)
    mw, OA, _ = harness.facade_arrays(scene, world, caps=caps, **kw)
    has_joints = len(rw.joints) > 0
    init_fields = INIT_FIELDS + (JOINT_INIT_FIELDS if has_joints else [])
    state_fields = harness.STATE_FIELDS + (harness.JOINT_STATE_FIELDS if has_joints else [])
    bad = harness.diff_state(RA, OA, fields=[f for f in init_fields if not f.startswith("s_")])
    if bad:
        raise SystemExit("%s: facade initial state differs from the reference: %s" % (case, bad))
    out = {"meta": np.array([dt, vi, pi, steps, world], dtype=np.float64),
           "caps": np.array([caps.nb, caps.nf, caps.nc, caps.npair, caps.nl] + ([caps.nj] if has_joints else []), dtype=np.int32),
           "scene": np.array(scene)}
    if kw:
        out["scene_kw"] = np.array(json.dumps(kw, sort_keys=True))      # scene keywords (read back by tests/conftest.py)
    for f in init_fields:
        out["init_" + f] = used(RA, f)
    n_sh = len(table)
    for f in ("s_meta", "s_radius", "s_data"):
        out["init_" + f] = np.array(RA[f][:n_sh], copy=True, order="C")
    orc = oracle.Oracle(OA, 0)
    rec = refpin.Recorder(rw)
    digests = np.zeros((steps, 32), np.uint8)
    ihash = np.zeros((steps, 32), np.uint8)
    ints = np.zeros((steps, 8), np.int32)
    t0 = time.time()
    for s in range(steps):
        rec.begin_step()
        rw.step(dt, vi, pi)
        orc.step(dt, vi, pi)
        refpin.extract(rw, RA, 0, table, schema)
        bad = harness.diff_state(RA, OA, fields=state_fields)
        if rec.islands != harness.oracle_islands(OA):
            bad.append("islands")
        if has_joints and rec.island_joints != harness.oracle_island_joints(OA):
            bad.append("island joints")
        if rec.pairs[-1] != harness.oracle_pairs(OA):
            bad.append("pairs")
        if bad:
            raise SystemExit("%s: oracle != reference at step %d: %s" % (case, s, bad))
        digests[s] = np.frombuffer(state_digest(OA), np.uint8)
        ihash[s] = np.frombuffer(int_hash(OA), np.uint8)
        ints[s] = int_summary(OA)
        if (s + 1) in cks:
            for f in CK_FIELDS + (JOINT_CK_FIELDS if has_joints else []):
                out["ck%d_%s" % (s + 1, f)] = used(OA, f)
    out["digest"] = digests
    out["ihash"] = ihash
    out["ints"] = ints
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, case + ".npz")
    np.savez_compressed(path, **out)
    print("%s: %d steps bit-identical to the reference (%.0f s), final ints %s -> %s (%d KB)" % (
        case, steps, time.time() - t0, ints[-1].tolist(), os.path.relpath(path, harness._ROOT),
        os.path.getsize(path) // 1024), flush=True)


if __name__ == "__main__":
    if not refpin.available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    for c in (sys.argv[1:] or list(CASES)):
        generate(c)
