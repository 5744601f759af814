"""Development probe: per-phase cycles and per-kernel ms during the landing transient (steps 10..60 from t=0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
w = scenes.pyramid(b2)
caps = schema.Caps(211, 211, 640, npair=1280)
for prof in ((False,) if os.environ.get("NO_PROF") else (False, True)):
    bw = b2.BatchedWorld.replicate(w, n, jitter=1e-2, config=3, real="f32", caps=caps, lanes_aux=int(os.environ.get("LANES_AUX", "0")))
    bw.step(1/60, 8, 3, 10)
    torch.cuda.synchronize()
    for lo in (10, 35, 60, 110):
        steps = 25 if lo < 60 else 50
        if prof:
            bw.step(1/60, 8, 3, steps, profile=True)
            p = bw.profile(); ws = max(p.pop("world_steps"), 1)
            keys = ("pairs0", "collide", "adjacency", "islands", "s_schedule", "sync", "pairs", "clear")
            print("steps %d..%d plan-side cycles/world-step:" % (lo, lo + steps), {k: int(p[k] / ws) for k in keys if k in p}, flush=True)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); bw.step(1/60, 8, 3, steps, time_kernels=True); e1.record(); torch.cuda.synchronize()
            k = bw.kernel_ms()
            print("steps %d..%d: %.2f ms/step; collide %.2f plan %.2f solve %.2f broad %.2f" % (lo, lo + steps, e0.elapsed_time(e1) / steps,
                  k["collide"] / steps, k["plan"] / steps, k["solve"] / steps, k["broadphase"] / steps), flush=True)
