"""Shared helpers for oracle validation and golden generation (TEST INFRASTRUCTURE ONLY)."""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for p in (_ROOT, _HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from pypybox2d_b200 import schema, scenes  # noqa: E402

STATE_FIELDS = ["w_count", "w_flags", "b_xf", "b_pos", "b_vel", "b_pos0", "b_force", "b_flags",
                "f_fat", "f_tight", "f_moved", "c_fix", "c_flags", "c_mat", "c_mtype", "c_mf", "c_pts", "c_imp",
                "c_id"]


def default_caps(nb, nf, nj=0):
    return schema.Caps(nb, nf, 4 * nb + 64, nj=nj)


def set_shapes(A, table):
    m, r, d = table.arrays()
    n = len(m)
    if n > A["s_meta"].shape[0]:
        raise ValueError("shape table capacity")
    A["s_meta"][:n] = m
    A["s_radius"][:n] = r
    A["s_data"][:n] = d


def facade_arrays(name, world=0, caps=None, n_shapes=512, **kw):
    """Scene built with THIS package's facade, packed into fp64 schema arrays (batch of one)."""
    import pypybox2d_b200 as mine
    mw = scenes.SCENES[name](mine, world=world, **kw)
    nb = len(mw.bodies)
    nf = sum(f._shape.child_count for b in mw.bodies for f in b.fixtures)      # proxies: one per shape child (Loop)
    caps = caps or default_caps(nb, nf, len(mw.joints))
    table = schema.ShapeTable()
    A = schema.alloc(caps, 1, n_shapes)
    schema.pack_world(mw, A, 0, table)
    set_shapes(A, table)
    return mw, A, caps


def reference_arrays(name, world=0, caps=None, n_shapes=512, **kw):
    """Scene built in the live reference and extracted into fp64 schema arrays (batch of one)."""
    import refpin
    b2 = refpin.load()
    rw = scenes.SCENES[name](b2, world=world, **kw)
    refpin.index_fixtures(rw)
    nb = len(rw.bodies)
    nf = sum(f._shape.child_count for b in rw.bodies for f in b.fixtures)
    caps = caps or default_caps(nb, nf, len(rw.joints))
    table = schema.ShapeTable()
    A = schema.alloc(caps, 1, n_shapes)
    refpin.extract(rw, A, 0, table, schema)
    set_shapes(A, table)
    return rw, A, caps, table


def used_view(A, name, w=0):
    """The rows of field `name` that are in use for world `w` (per w_count)."""
    nb, nf, nc, npair = [int(x) for x in A["w_count"][w]]
    scope = {f[0]: f[1] for f in schema.FIELDS}[name]
    a = A[name][w] if scope != "S" else A[name]
    if scope == "B":
        return a[:nb]
    if scope == "F":
        return a[:nf]
    if scope == "C":
        return a[:nc]
    if scope == "J":
        return a[:int(A["w_joint"][w][0])]
    return a


def diff_state(A, B, w=0, fields=STATE_FIELDS, exact=True, rtol=0.0, atol=0.0):
    """Names of fields whose used rows differ between two array dicts."""
    bad = []
    for k in fields:
        x, y = used_view(A, k, w), used_view(B, k, w)
        if k == "w_count":
            x, y = x[:3], y[:3]
        if k == "w_joint":
            x, y = x[:1], y[:1]
        if x.shape != y.shape:
            bad.append(k)
            continue
        if exact or x.dtype.kind in "iu":
            mask = 0xFFFFFFFF
            if k == "b_flags":      # island flag is transient bookkeeping
                x = x & ~np.uint32(schema.BODY_ISLAND)
                y = y & ~np.uint32(schema.BODY_ISLAND)
            if k == "w_flags":      # device-only bookkeeping bit
                x = x & ~np.uint32(schema.WORLD_GRAPH_VALID | schema.WORLD_TOI_DIRTY)
                y = y & ~np.uint32(schema.WORLD_GRAPH_VALID | schema.WORLD_TOI_DIRTY)
            if k == "c_flags":
                x = x & ~np.uint32(schema.CONTACT_ISLAND)
                y = y & ~np.uint32(schema.CONTACT_ISLAND)
            if k == "j_def":        # island flag is transient bookkeeping
                x = x & ~np.int32(schema.JOINT_ISLAND)
                y = y & ~np.int32(schema.JOINT_ISLAND)
            if not np.array_equal(x, y):
                bad.append(k)
        else:
            if not np.allclose(x, y, rtol=rtol, atol=atol):
                bad.append(k)
    return bad


JOINT_STATE_FIELDS = ["w_joint", "j_def", "j_par", "j_imp", "j_adj", "b_jadj"]


def oracle_island_joints(A, w=0):
    """[joint list] per island of the last step."""
    n_isl = int(A["w_island"][w][0])
    nlj = int(A["w_joint"][w][1])
    st = A["l_jstart"][w]
    out = []
    for i in range(n_isl):
        j0 = int(st[i])
        j1 = int(st[i + 1]) if i + 1 < n_isl else nlj
        out.append([int(v) for v in A["l_joint"][w][j0:j1]])
    return out


def oracle_pairs(A, w=0):
    n = int(A["w_count"][w][3])
    return [tuple(int(v) for v in p) for p in A["p_pairs"][w][:n]]


def oracle_islands(A, w=0):
    """[(body list, contact list)] of the last step, from the observable scratch fields."""
    n_isl, nlb, nlc, _ = [int(x) for x in A["w_island"][w]]
    out = []
    st = A["l_start"][w]
    for i in range(n_isl):
        b0, c0 = int(st[i][0]), int(st[i][1])
        b1, c1 = (int(st[i + 1][0]), int(st[i + 1][1])) if i + 1 < n_isl else (nlb, nlc)
        out.append(([int(v) for v in A["l_body"][w][b0:b1]], [int(v) for v in A["l_contact"][w][c0:c1]]))
    return out
