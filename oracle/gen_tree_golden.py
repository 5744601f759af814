#!/usr/bin/env python
"""Generate tests/golden_queries/tree.npz from the LIVE reference (build container only): the broadphase tree's topology and
its observables (`tree_height`, `tree_balance`, `tree_quality`; broadphase.py:137-150 over
dynamictree.py).

Two parts:
  ops_*   a seeded random sequence of create / move / destroy operations driven through the reference's `DynamicTree`
          (create_proxy, move_proxy, destroy_proxy: dynamictree.py:115-188); after every operation the tree's height, area
          ratio, the largest sibling height difference and the preorder of the tree (inner node -1, leaf = proxy number).
          tests/test_tree.py replays the sequence through `pypybox2d_b200.proxytree.ProxyTree` on the CPU.
  world_* the mutation scenario of gen_mutation_golden.py (an 8-row pyramid with bodies and fixtures destroyed and created
          between steps) stepped by the reference; after every step the world's tree_height, tree_quality and the preorder
          with leaves named by fixture index.  The GPU test plays the scenario through `pypybox2d_b200.World`.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa: E402
import refpin  # noqa: E402
import gen_mutation_golden as gm  # noqa: E402

N_OPS = 1500
SEED = 20261018


def make_ops(rng, n_ops):
    """[(kind, proxy, lx, ly, ux, uy, dx, dy)]: kind 0 create, 1 move, 2 destroy.  Boxes cluster in a 40 x 20 region so that
    moves mostly stay inside the fat box (no re-insertion) and sometimes leave it."""
    ops, live, boxes, nxt = [], [], {}, 0
    for _ in range(n_ops):
        r = rng.random()
        if not live or r < 0.25 or (len(live) < 8 and r < 0.6):
            c = rng.uniform((-20, 0), (20, 20))
            h = rng.uniform(0.1, 1.5, 2)
            box = (c[0] - h[0], c[1] - h[1], c[0] + h[0], c[1] + h[1])
            ops.append((0, nxt) + box + (0.0, 0.0))
            live.append(nxt); boxes[nxt] = box; nxt += 1
        elif r < 0.85:
            p = live[int(rng.integers(len(live)))]
            d = rng.normal(0.0, 0.15, 2) * (8.0 if rng.random() < 0.1 else 1.0)
            b = boxes[p]
            box = (b[0] + d[0], b[1] + d[1], b[2] + d[0], b[3] + d[1])
            ops.append((1, p) + box + (float(d[0]), float(d[1])))
            boxes[p] = box
        else:
            p = live.pop(int(rng.integers(len(live))))
            ops.append((2, p, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0))
    return np.array(ops, np.float64)


def preorder(root, name_of):
    out, stack = [], [root] if root is not None else []
    while stack:
        n = stack.pop()
        if n.child1 is None:
            out.append(name_of(n))
        else:
            out.append(-1)
            stack.append(n.child2)
            stack.append(n.child1)
    return out


def sibling_height_difference(tree):
    return max([abs(n.child2.height - n.child1.height) for n in tree._nodes if n.height > 1] or [0])


def main():
    b2 = refpin.load()
    from pypybox2d.dynamictree import DynamicTree
    from pypybox2d.common import AABB, Vec2
    rng = np.random.default_rng(SEED)
    ops = make_ops(rng, N_OPS)
    tree, nodes, names = DynamicTree(), {}, {}
    obs, pre, moved = [], [], []

    class Tag(object):
        pass
    for kind, p, lx, ly, ux, uy, dx, dy in ops:
        kind, p = int(kind), int(p)
        m = 0
        if kind == 0:
            tag = Tag()
            names[p] = tag
            nodes[p] = tree.create_proxy(AABB(Vec2(lx, ly), Vec2(ux, uy)), tag)
            tag.p = p
        elif kind == 1:
            m = 1 if tree.move_proxy(nodes[p], AABB(Vec2(lx, ly), Vec2(ux, uy)), Vec2(dx, dy)) else 0
        else:
            tree.destroy_proxy(nodes.pop(p))
        obs.append((tree.height, tree.area_ratio, sibling_height_difference(tree)))
        pre.append(np.array(preorder(tree._root, lambda n: n.user_data.p), np.int32))
        moved.append(m)
    assert tree.max_balance is None          # dynamictree.py:285-297 has no `return`
    out = {"ops": ops, "ops_obs": np.array(obs, np.float64), "ops_moved": np.array(moved, np.uint8),
           "ops_pre_len": np.array([len(x) for x in pre], np.int32), "ops_pre": np.concatenate(pre)}

    from pypybox2d_b200 import scenes
    world = scenes.pyramid(b2, world=1, rows=8)
    refpin.index_fixtures(world)
    wobs, wpre = [], []

    def record(w):
        bp = w.contact_manager.broadphase
        fidx = {id(f): j for j, f in enumerate([f for b in w.bodies for f in b.fixtures])}
        wobs.append((bp.tree_height, bp.tree_quality, sibling_height_difference(bp._tree), bp.proxy_count))
        wpre.append(np.array(preorder(bp._tree._root, lambda n: fidx[id(n.user_data.fixture)]), np.int32))
        assert bp.tree_balance is None
    gm.script(b2, world, record, reindex=refpin.index_fixtures)
    out.update({"world_obs": np.array(wobs, np.float64), "world_pre_len": np.array([len(x) for x in wpre], np.int32),
                "world_pre": np.concatenate(wpre)})
    path = os.path.join(harness._ROOT, "tests", "golden_queries", "tree.npz")
    np.savez_compressed(path, **out)
    h = out["ops_obs"][:, 0]
    print("ops: %d (re-inserted moves %d), height up to %d; world: %d steps, height %d..%d, quality %.3f..%.3f -> %s (%d KB)" % (
        len(ops), int(sum(moved)), int(h.max()), len(wobs), int(out["world_obs"][:, 0].min()), int(out["world_obs"][:, 0].max()),
        out["world_obs"][:, 1].min(), out["world_obs"][:, 1].max(), os.path.relpath(path, harness._ROOT), os.path.getsize(path) // 1024))


if __name__ == "__main__":
    if not refpin.available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    main()
