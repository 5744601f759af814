"""Small run for compute-sanitizer over the round-2 kernels: the TOI kernel (12 jittered bullet worlds, continuous collision on;
12 landing pyramids with it on) and the query kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for name, cfg, kw in (("bullets", 7, {}), ("pyramid", 3, {"continuous_physics": True})):
    bw = b2.BatchedWorld.replicate(scenes.SCENES[name](b2, **kw), 12, jitter=1e-2, config=cfg, real="f32")
    bw.step(1 / 60, 8, 3, steps)
    torch.cuda.synchronize()
    print(name, "status", bw.status(False), "contacts", int(bw.t["w_count"][:, 2].max()))
    rng = np.random.default_rng(1)
    n = 64
    w = rng.integers(0, 12, n)
    c = rng.uniform(-10, 10, (n, 2)); h = rng.uniform(0.1, 4, (n, 2))
    cnt, fix = bw.query_aabb(w, np.concatenate([c - h, c + h], 1), max_hits=8)
    seg = rng.uniform(-15, 15, (n, 4))
    rc, rf, rh = bw.ray_cast(w, seg, max_hits=8)
    r1, _, _ = bw.ray_cast(w, seg, closest=True, max_hits=1)
    ins = bw.test_points(w, rng.integers(0, 20, n), rng.uniform(-5, 5, (n, 2)))
    print(name, "queries", int(cnt.sum()), int(rc.sum()), int(r1.sum()), int(ins.sum()))
