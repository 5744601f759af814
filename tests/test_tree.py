"""The reference's broadphase tree topology and its observables (`tree_height`, `tree_balance`,
`tree_quality`; broadphase.py:137-150).  tests/golden_queries/tree.npz holds the LIVE reference's answers
(oracle/gen_tree_golden.py): a seeded random create / move / destroy sequence through its `DynamicTree`, and its tree after every
step of the host-mutation scenario.  `ProxyTree` must reproduce both exactly: the same topology (preorder of inner nodes and
leaves), the same height, and the same area ratio to the bit (it sums node perimeters in the reference's pool order)."""
import os

import numpy as np
import pytest

import gen_mutation_golden as gm
from pypybox2d_b200 import scenes
from pypybox2d_b200.proxytree import ProxyTree, NIL
from pypybox2d_b200.settings import AABB_EXTENSION, AABB_MULTIPLIER

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_queries", "tree.npz")


def preorder(tree, name_of):
    out, stack = [], [tree.root] if tree.root != NIL else []
    while stack:
        n = stack.pop()
        if tree.child1[n] == NIL:
            out.append(name_of(n))
        else:
            out.append(-1)
            stack.append(tree.child2[n])
            stack.append(tree.child1[n])
    return out


def test_random_operations_reproduce_the_reference_tree():
    G = np.load(GOLDEN)
    ends = np.cumsum(G["ops_pre_len"])
    tree, slot, name = ProxyTree(), {}, {}
    for k, (kind, p, lx, ly, ux, uy, dx, dy) in enumerate(G["ops"]):
        kind, p = int(kind), int(p)
        moved = 0
        if kind == 0:
            s = tree.create_leaf((lx - AABB_EXTENSION, ly - AABB_EXTENSION, ux + AABB_EXTENSION, uy + AABB_EXTENSION))
            slot[p] = s
            name[s] = p
        elif kind == 1:
            moved = 1 if tree.move_leaf(slot[p], (lx, ly, ux, uy), (dx, dy), AABB_EXTENSION, AABB_MULTIPLIER) else 0
        else:
            name.pop(slot[p])
            tree.destroy_leaf(slot.pop(p))
        height, ratio, diff = G["ops_obs"][k]
        assert moved == int(G["ops_moved"][k]), "operation %d: re-insertion decision" % k
        assert tree.tree_height == int(height), "operation %d: height" % k
        assert tree.tree_quality == ratio, "operation %d: area ratio %r != %r" % (k, tree.tree_quality, ratio)
        assert tree.max_balance == int(diff), "operation %d: sibling height difference" % k
        want = G["ops_pre"][ends[k] - G["ops_pre_len"][k]:ends[k]]
        assert preorder(tree, lambda n: name[n]) == list(want), "operation %d: topology" % k
    assert tree.tree_balance is None           # the reference's max_balance has no `return` (dynamictree.py:285-297)
    assert tree.count == 2 * len(slot) - 1     # a full binary tree over the live leaves


def test_empty_and_single_leaf_tree():
    t = ProxyTree()
    assert t.tree_height == 0 and t.tree_quality == 0.0 and t.tree_balance is None and t.structure() is None
    a = t.create_leaf((0.0, 0.0, 1.0, 1.0))
    assert t.tree_height == 0 and t.tree_quality == 1.0 and t.structure() == a
    t.destroy_leaf(a)
    assert t.root == NIL and t.count == 0
    for i in range(40):                        # the pool doubles (16 -> 32 -> 64 -> 128) and keeps the reference's slot order
        t.create_leaf((i, 0.0, i + 0.5, 0.5))
    assert len(t.parent) == 128 and t.count == 79


def test_facade_tracks_proxies_without_a_device():
    import pypybox2d_b200 as b2
    w = b2.World(continuous_physics=False)
    bp = w.contact_manager.broadphase
    assert bp.proxy_count == 0 and bp.tree_height == 0 and bp.tree_quality == 0.0
    g = w.create_static_body()
    g.create_edge_fixture((-10, 0), (10, 0))
    a = w.create_dynamic_body(position=(0, 1), fixtures=b2.Fixture(b2.Circle(0.5), density=1))
    f = a.create_circle_fixture(0.2, (0.5, 0), density=1.0)
    assert bp.proxy_count == 3 and bp.tree_height == 2 and bp.tree_balance is None
    a.destroy_fixture(f)
    assert bp.proxy_count == 2 and bp.tree_height == 1
    w.destroy_body(a)
    assert len(w._tree_leaf) == 1 and bp.tree_height == 0 and bp.tree_quality == 1.0
    assert bp.proxy_count == 0                          # the reference's counter loses two per proxy of a destroyed body (body.py:330-340)
    off = w.create_dynamic_body(position=(3, 3), active=False, fixtures=b2.Fixture(b2.Circle(0.5), density=1))
    assert bp.proxy_count == 0 and len(w._tree_leaf) == 1 and off in w.bodies      # an inactive body's fixtures have no proxies (body.py:292-297)
    off.active = True                                   # body.py:858-862: proxies created
    assert bp.proxy_count == 1 and len(w._tree_leaf) == 2 and bp.tree_height == 1
    off.active = False                                  # body.py:863-866
    assert bp.proxy_count == 0 and len(w._tree_leaf) == 1 and bp.tree_height == 0


@pytest.mark.gpu
def test_tree_observables_follow_the_reference_through_the_mutation_script():
    import pypybox2d_b200 as b2
    G = np.load(GOLDEN)
    ends = np.cumsum(G["world_pre_len"])
    world = scenes.pyramid(b2, world=1, rows=8, real="f64")
    bp = world.contact_manager.broadphase
    step = [0]
    worst = [0.0]

    def record(w):
        i = step[0]
        height, quality, diff, count = G["world_obs"][i]
        fidx = {id(f): j for j, f in enumerate([f for b in w.bodies for f in b.fixtures])}
        name = {slot: fidx[key[0]] for key, slot in w._tree_leaf.items()}
        want = G["world_pre"][ends[i] - G["world_pre_len"][i]:ends[i]]
        assert bp.proxy_count == int(count), "proxy count at step %d" % i
        assert preorder(w._tree, lambda n: name[n]) == list(want), "tree topology at step %d" % i
        assert bp.tree_height == int(height), "tree height at step %d" % i
        assert w._tree.max_balance == int(diff)
        rel = abs(bp.tree_quality - quality) / quality
        worst[0] = max(worst[0], rel)
        assert rel <= 1e-9, "tree quality at step %d: %r against %r" % (i, bp.tree_quality, quality)
        assert bp.tree_balance is None
        step[0] += 1
    gm.script(b2, world, record)
    assert step[0] == len(G["world_obs"])
    print("tree observables: %d steps, topology identical, worst relative quality difference %.2e" % (step[0], worst[0]))
