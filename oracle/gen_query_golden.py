#!/usr/bin/env python
"""Generate tests/golden/queries.npz from the LIVE reference (build container only): World.query_aabb, World.ray_cast and
Fixture.test_point (world.py:371-440, shapes.py) on four scenes stepped to a busy moment, for seeded random queries.

For every scene the reference and the C oracle are stepped side by side (state asserted identical), then every query is
answered by both; the script FAILS unless the fixture SETS of query_aabb are equal, every ray's hits (fixture, fraction,
normal, point -- bitwise) are equal, the closest hit found through the reference's own clipping protocol is the
smallest-fraction hit of that set, and test_point agrees.  The fixture stores the queries and the answers; tests replay
them on the oracle (CPU) and on the device kernels (b2b_query_aabb / b2b_ray_cast / b2b_test_points).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa: E402
import oracle  # noqa: E402
import refpin  # noqa: E402
from pypybox2d_b200 import schema  # noqa: E402

# scene, world, dt, vel_iters, pos_iters, steps, scene keywords, query window (lx, ly, ux, uy)
CASES = [("features", 1, 1.0 / 60, 8, 3, 80, {}, (-22.0, -1.0, 22.0, 16.0)),
         ("mixed", 1, 1.0 / 60, 8, 3, 120, {}, (-11.0, -1.0, 11.0, 21.0)),
         ("bullets", 1, 1.0 / 60, 8, 3, 40, {}, (-21.0, -1.0, 21.0, 13.0)),
         ("pyramid", 1, 1.0 / 60, 8, 3, 100, {}, (-10.0, -1.0, 18.0, 27.0))]
N_BOX, N_RAY, N_PT, MAX_HITS = 96, 160, 256, 64


def main():
    b2 = refpin.load()
    from pypybox2d.common import AABB
    out = {"n_cases": np.array(len(CASES)), "max_hits": np.array(MAX_HITS)}
    for k, (scene, world, dt, vi, pi, steps, kw, win) in enumerate(CASES):
        rw, RA, caps, table = harness.reference_arrays(scene, world, **kw)
        mw, OA, _ = harness.facade_arrays(scene, world, caps=caps, **kw)
        orc = oracle.Oracle(OA, 0)
        for s in range(steps):
            rw.step(dt, vi, pi)
            orc.step(dt, vi, pi)
        refpin.extract(rw, RA, 0, table, schema)
        bad = harness.diff_state(RA, OA)
        if bad:
            raise SystemExit("%s: oracle != reference after %d steps: %s" % (scene, steps, bad))
        fixtures = [f for b in rw.bodies for f in b.fixtures]
        fidx = {id(f): j for j, f in enumerate(fixtures)}
        rng = np.random.default_rng([20261018, k])
        lx, ly, ux, uy = win
        # boxes: small, medium and a few huge ones
        c = rng.uniform((lx, ly), (ux, uy), size=(N_BOX, 2))
        h = rng.uniform(0.05, 3.0, size=(N_BOX, 2)) * rng.choice([0.2, 1.0, 4.0], size=(N_BOX, 1))
        boxes = np.concatenate([c - h, c + h], axis=1)
        box_count = np.zeros(N_BOX, np.int32)
        box_hits = np.full((N_BOX, MAX_HITS), -1, np.int32)
        for i, bx in enumerate(boxes):
            ref = sorted(fidx[id(f)] for f in rw.query_aabb(AABB(lower_bound=(bx[0], bx[1]), upper_bound=(bx[2], bx[3]))))
            got = orc.query_aabb(bx).tolist()
            if ref != got:
                raise SystemExit("%s: query_aabb %d differs: %s vs %s" % (scene, i, ref, got))
            box_count[i] = len(ref)
            box_hits[i, :min(len(ref), MAX_HITS)] = ref[:MAX_HITS]
        # rays: random segments, plus rays aimed at body centres so that most of them hit something
        p1 = rng.uniform((lx, ly), (ux, uy), size=(N_RAY, 2))
        p2 = rng.uniform((lx, ly), (ux, uy), size=(N_RAY, 2))
        nb = len(rw.bodies)
        for i in range(0, N_RAY, 2):
            b = rw.bodies[int(rng.integers(0, nb))]
            d = np.array([b.position.x, b.position.y]) - p1[i]
            p2[i] = p1[i] + d * float(rng.uniform(1.05, 2.5))
        rays = np.concatenate([p1, p2], axis=1)
        ray_count = np.zeros(N_RAY, np.int32)
        ray_fix = np.full((N_RAY, MAX_HITS), -1, np.int32)
        ray_hit = np.zeros((N_RAY, MAX_HITS, 5))
        closest = np.full(N_RAY, -1, np.int32)
        for i, r in enumerate(rays):
            ref = {}
            for ret in rw.ray_cast((r[0], r[1]), (r[2], r[3])):
                fixture, point, normal, fraction = ret
                ref[fidx[id(fixture)]] = (fraction, normal.x, normal.y, point.x, point.y)
                ret[-1] = 1.0                                   # do not clip: every fixture on the ray is reported
            gf, gh = orc.ray_cast(r[:2], r[2:])
            got = {int(f): tuple(float(x) for x in row) for f, row in zip(gf, gh)}
            if ref != got:
                raise SystemExit("%s: ray_cast %d differs:\n ref %s\n got %s" % (scene, i, ref, got))
            # the reference's own closest-hit protocol (the generator's default: clip to each reported fraction)
            best = None
            for ret in rw.ray_cast((r[0], r[1]), (r[2], r[3])):
                fj = fidx[id(ret[0])]
                if best is None or ret[3] < best[0] or (ret[3] == best[0] and fj < best[1]):
                    best = (ret[3], fj)
            if ref:
                want = min((v[0], f) for f, v in ref.items())
                if best is None or best != want:
                    raise SystemExit("%s: closest hit of ray %d: protocol %s, set %s" % (scene, i, best, want))
                closest[i] = want[1]
            elif best is not None:
                raise SystemExit("%s: ray %d: protocol found %s, set empty" % (scene, i, best))
            fs = sorted(ref)
            ray_count[i] = len(fs)
            for j, f in enumerate(fs[:MAX_HITS]):
                ray_fix[i, j] = f
                ray_hit[i, j] = ref[f]
        # points: around fixtures' bodies so that a good share is inside
        pf = rng.integers(0, len(fixtures), size=N_PT).astype(np.int32)
        pts = np.zeros((N_PT, 2))
        inside = np.zeros(N_PT, np.uint8)
        for i, f in enumerate(pf):
            fx = fixtures[int(f)]
            c0 = fx._body.position
            pts[i] = (c0.x + float(rng.normal(0, 0.45)), c0.y + float(rng.normal(0, 0.45)))
            try:
                ref = bool(fx.test_point(b2.Vec2(pts[i, 0], pts[i, 1])))
            except NotImplementedError:
                ref = False
            except Exception:
                ref = False                                     # edges have no interior (Shape.test_point is not defined for them)
            got = orc.test_point(int(f), pts[i])
            if ref != got:
                raise SystemExit("%s: test_point %d (fixture %d) differs: %s vs %s" % (scene, i, f, ref, got))
            inside[i] = ref
        pre = "q%d_" % k
        out[pre + "scene"] = np.array(scene)
        out[pre + "meta"] = np.array([dt, vi, pi, steps, world], np.float64)
        out[pre + "caps"] = np.array([caps.nb, caps.nf, caps.nc, caps.npair, caps.nl], np.int32)
        out[pre + "boxes"], out[pre + "box_count"], out[pre + "box_hits"] = boxes, box_count, box_hits
        out[pre + "rays"], out[pre + "ray_count"], out[pre + "ray_fix"], out[pre + "ray_hit"] = rays, ray_count, ray_fix, ray_hit
        out[pre + "closest"] = closest
        out[pre + "pt_fix"], out[pre + "pts"], out[pre + "inside"] = pf, pts, inside
        print("%s: %d boxes (%d overlaps), %d rays (%d hit, %d hits), %d points (%d inside): oracle == reference" % (
            scene, N_BOX, int(box_count.sum()), N_RAY, int((ray_count > 0).sum()), int(ray_count.sum()), N_PT, int(inside.sum())),
            flush=True)
    path = os.path.join(harness._ROOT, "tests", "golden_queries", "queries.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print("->", os.path.relpath(path, harness._ROOT), os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    if not refpin.available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    main()
