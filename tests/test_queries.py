"""Queries: World.query_aabb, World.ray_cast, Fixture.test_point.

tests/golden_queries/queries.npz was written by oracle/gen_query_golden.py from the LIVE reference: seeded random boxes, rays
and points on four scenes, with the reference's answers (and the generator's proof that the C oracle gives the same ones).
CPU: the oracle replays them.  GPU: the device kernels (b2b_query_aabb / b2b_ray_cast / b2b_test_points) answer them through
the C-ABI, in fp64 against the stored answers and in fp32 against the oracle on the fp32-rounded state.
"""
import os

import numpy as np
import pytest

import harness
import oracle
from pypybox2d_b200 import schema

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_queries", "queries.npz"))
N_CASES = int(G["n_cases"])
MAX_HITS = int(G["max_hits"])


def stepped_oracle(k):
    pre = "q%d_" % k
    dt, vi, pi, steps, world = G[pre + "meta"]
    caps = schema.Caps(*[int(x) for x in G[pre + "caps"]])
    mw, A, _ = harness.facade_arrays(str(G[pre + "scene"]), int(world), caps=caps)
    orc = oracle.Oracle(A, 0)
    orc.step(float(dt), int(vi), int(pi), int(steps))
    return A, orc, caps, pre


@pytest.mark.parametrize("k", range(N_CASES))
def test_oracle_answers_the_reference_queries(k):
    A, orc, caps, pre = stepped_oracle(k)
    for i, box in enumerate(G[pre + "boxes"]):
        got = orc.query_aabb(box)
        assert len(got) == int(G[pre + "box_count"][i])
        assert got[:MAX_HITS].tolist() == [f for f in G[pre + "box_hits"][i].tolist() if f >= 0]
    for i, ray in enumerate(G[pre + "rays"]):
        fix, hit = orc.ray_cast(ray[:2], ray[2:])
        n = int(G[pre + "ray_count"][i])
        assert len(fix) == n
        m = min(n, MAX_HITS)
        assert fix[:m].tolist() == G[pre + "ray_fix"][i][:m].tolist()
        assert np.array_equal(hit[:m], G[pre + "ray_hit"][i][:m])              # bitwise: fraction, normal, point
        if n:
            best = min((hit[j][0], int(fix[j])) for j in range(n))
            assert best[1] == int(G[pre + "closest"][i])
    for i in range(len(G[pre + "pts"])):
        assert orc.test_point(int(G[pre + "pt_fix"][i]), G[pre + "pts"][i]) == bool(G[pre + "inside"][i])


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(N_CASES))
def test_device_queries_fp64_match_the_reference(k):
    import pypybox2d_b200 as b2
    A, orc, caps, pre = stepped_oracle(k)
    # a batch of three copies of the world: every query goes to world 1, worlds 0 and 2 must not leak into the answers
    B = {name: (np.repeat(v, 3, axis=0) if not name.startswith("s_") else v.copy()) for name, v in A.items()}
    B["b_xf"][0, :, 0] += 1000.0
    B["f_fat"][2, :, :] += 1000.0
    bw = b2.BatchedWorld(B, caps, real="f64")
    boxes, rays = G[pre + "boxes"], G[pre + "rays"]
    cnt, fix = bw.query_aabb(np.ones(len(boxes), np.int32), boxes, max_hits=MAX_HITS)
    assert np.array_equal(cnt, G[pre + "box_count"])
    assert np.array_equal(fix, G[pre + "box_hits"])
    cnt, fix, hit = bw.ray_cast(np.ones(len(rays), np.int32), rays, max_hits=MAX_HITS)
    assert np.array_equal(cnt, G[pre + "ray_count"])
    assert np.array_equal(fix, G[pre + "ray_fix"])
    m = G[pre + "ray_fix"] >= 0
    assert np.abs(hit[m] - G[pre + "ray_hit"][m]).max() <= 1e-12                 # (radius ** 2 is libm pow on the CPU, a product here)
    cnt1, fix1, hit1 = bw.ray_cast(np.ones(len(rays), np.int32), rays, closest=True, max_hits=1)
    assert np.array_equal(cnt1, (G[pre + "ray_count"] > 0).astype(np.int32))
    assert np.array_equal(fix1[:, 0], G[pre + "closest"])
    for i in np.nonzero(cnt1)[0]:
        j = G[pre + "ray_fix"][i].tolist().index(int(fix1[i, 0]))
        assert np.abs(hit1[i, 0] - G[pre + "ray_hit"][i][j]).max() <= 1e-12
    inside = bw.test_points(np.ones(len(G[pre + "pts"]), np.int32), G[pre + "pt_fix"], G[pre + "pts"])
    assert np.array_equal(inside, G[pre + "inside"].astype(bool))
    # a zero-length ray is the reference's ValueError
    c0, _, _ = bw.ray_cast([1], [(1.0, 2.0, 1.0, 2.0)], max_hits=4)
    assert int(c0[0]) == -1


@pytest.mark.gpu
def test_device_queries_fp32_and_the_facade_protocol():
    import pypybox2d_b200 as b2
    from pypybox2d_b200 import scenes
    # fp32 kernels against the oracle on the fp32-rounded state: the same answers except where a query grazes a box or a shape
    A, orc, caps, pre = stepped_oracle(2)
    R = {k: (v.astype(np.float32).astype(np.float64) if v.dtype == np.float64 else v.copy()) for k, v in A.items()}
    o32 = oracle.Oracle(R, 0)
    bw = b2.BatchedWorld({k: v.copy() for k, v in R.items()}, caps, real="f32")
    boxes = G[pre + "boxes"].astype(np.float32).astype(np.float64)
    rays = G[pre + "rays"].astype(np.float32).astype(np.float64)
    cnt, fix = bw.query_aabb(np.zeros(len(boxes), np.int32), boxes, max_hits=MAX_HITS)
    same = sum(fix[i][:cnt[i]].tolist() == o32.query_aabb(boxes[i]).tolist() for i in range(len(boxes)))
    assert same >= len(boxes) - 2
    cnt, fix, hit = bw.ray_cast(np.zeros(len(rays), np.int32), rays, max_hits=MAX_HITS)
    agree, worst = 0, 0.0
    for i in range(len(rays)):
        of, oh = o32.ray_cast(rays[i][:2], rays[i][2:])
        if fix[i][:cnt[i]].tolist() == of.tolist():
            agree += 1
            if len(of):
                worst = max(worst, float(np.abs(hit[i][:len(of)] - oh).max()))
    assert agree >= len(rays) - 4 and worst <= 5e-3, (agree, worst)    # (rays up to 100 m long: a grazing circle hit moves by 1.7e-3 m in fp32)
    # the facade: the reference's generator / callback protocol on a world stepped through World.step
    world = scenes.bullets(b2, world=1, real="f64")
    for _ in range(40):
        world.step(1 / 60, 8, 3)
    fixtures = [f for b in world.bodies for f in b.fixtures]
    ray = G[pre + "rays"][0]
    want = [fixtures[j] for j in G[pre + "ray_fix"][0] if j >= 0]
    got_all = []
    for ret in world.ray_cast(ray[:2], ray[2:]):
        got_all.append(ret[0])
        ret[-1] = 1.0                                   # never clip: every fixture on the ray
    assert got_all == want
    seen = []
    world.ray_cast(ray[:2], ray[2:], callback=lambda f, p, n, fr: seen.append((fr, f)) or fr)     # clip to every hit
    if want:
        assert min(seen, key=lambda t: t[0])[1] is fixtures[int(G[pre + "closest"][0])]
    box = G[pre + "boxes"][3]
    hits = list(world.query_aabb(((box[0], box[1]), (box[2], box[3]))))
    assert hits == [fixtures[j] for j in G[pre + "box_hits"][3] if j >= 0]
    first = []
    world.query_aabb(((-100, -100), (100, 100)), callback=lambda f: first.append(f) and False)
    assert len(first) == 1
    k = int(np.nonzero(G[pre + "inside"])[0][0])
    assert fixtures[int(G[pre + "pt_fix"][k])].test_point(G[pre + "pts"][k])
    with pytest.raises(ValueError):
        list(world.ray_cast((0, 0), (0, 0)))
