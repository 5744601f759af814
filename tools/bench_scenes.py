"""Throughput of the OTHER BASELINE.json configurations on one GPU (the headline bench.py stays on the pyramid batch):
device-resident state, CUDA events, settled regime (distinct jittered worlds settled, then tiled) and the whole job from the
scripted start.  usage: bench_scenes.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
if os.environ.get("B2B_VARIANT"):      # development: an experimental build (tools/build_variant.sh)
    from pypybox2d_b200 import _lib
    _lib.SO_PATH = os.path.join(_lib.CSRC, "variants", "libb2b_%s.so" % os.environ["B2B_VARIANT"])
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes

CASES = [   # scene, config id, worlds on one GPU, dt, vel, pos, distinct, settle steps, what
    ("hello", 2, 65536, 0.1, 10, 10, 1024, 60, "configs[1] examples/hello-bench.py x 65,536 worlds"),
    ("vertical_stack", 2, 65536, 1 / 60, 10, 10, 2048, 300, "configs[1] examples/verticalstack-bench.py x 65,536 worlds"),
    ("mixed", 4, 131072, 1 / 60, 8, 3, 2048, 300, "configs[3] mixed circles + convex polygons, 256 bodies/world, 131,072 worlds"),
    ("deep_stacks", 5, 32768, 1 / 60, 8, 3, 2048, 300, "configs[4] 10 columns x 20 boxes, 32,768 worlds per GPU (262,144 on 8)"),
]
out = []
for name, cfg, n, dt, vi, pi, distinct, settle, what in CASES:
    w = scenes.SCENES[name](b2)
    n_dyn = sum(1 for b in w.bodies if b.type != 0) if hasattr(w.bodies[0], "type") else None
    seed = b2.BatchedWorld.replicate(w, distinct, jitter=1e-2, config=cfg, real="f32")
    if n_dyn is None:
        n_dyn = int(((seed.t["b_flags"][0] & 3) != 0).sum())
    # whole job from the scripted start on the seed batch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); seed.step(dt, vi, pi, settle); e1.record(); torch.cuda.synchronize()
    t0_ms = e0.elapsed_time(e1)
    assert seed.status(raise_on_overflow=False) == 0, name
    big = seed.tiled(n)
    del seed
    big.step(dt, vi, pi, 3)
    torch.cuda.synchronize(); e0.record(); big.step(dt, vi, pi, 10, time_kernels=True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    k = big.kernel_ms()
    row = {"scene": name, "what": what, "worlds": n, "dynamic_bodies_per_world": n_dyn, "step": [dt, vi, pi],
           "settled": {"ms_per_step": ms, "body_steps_per_s": n_dyn * n / (ms * 1e-3),
                       "kernel_ms": {q: k[q] / 10 for q in ("collide", "plan", "solve", "broadphase")},
                       "contacts_per_world": float(big.t["w_count"][:4096, 2].float().mean()),
                       "touching_per_world": float(big.t["w_island"][:4096, 2].float().mean())},
           "from_t0": {"worlds": distinct, "steps": settle, "ms_per_step": t0_ms / settle,
                       "body_steps_per_s": n_dyn * distinct * settle / (t0_ms * 1e-3)},
           "status": int(big.status(raise_on_overflow=False))}
    print(json.dumps(row), flush=True)
    out.append(row)
    del big
    torch.cuda.empty_cache()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
