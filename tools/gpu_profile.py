"""Development probe: per-phase cycle breakdown of the step kernel on settled pyramid worlds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema

def run(n_worlds, real, threads, settle=200, steps=4, ctas=0):
    w = scenes.pyramid(b2)
    caps = schema.Caps(211, 211, 640, npair=1280)
    bw = b2.BatchedWorld.replicate(w, n_worlds, jitter=1e-2, config=3, real=real, caps=caps, threads=threads, ctas_per_sm=ctas)
    bw.step(1/60, 8, 3, settle)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bw.step(1/60, 8, 3, steps, profile=True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    p = bw.profile()
    ws = max(p.pop("world_steps"), 1)
    tot = sum(p[k] for k in ("pairs0","collide","adjacency","islands","solve","sync","pairs","clear"))
    info = bw.info()
    print("%s worlds=%d threads=%d grid=%d smem=%d: %.2f ms/step %.3e body-steps/s; cycles/world-step %.0f status=%d" % (
        real, n_worlds, threads, info["grid"], info["smem_bytes"], ms/steps, 210*n_worlds*steps/(ms*1e-3), tot/ws, bw.status(False)))
    print("   top:", {k: "%.1f%%" % (100*p[k]/tot) for k in ("collide","adjacency","islands","solve","sync","pairs","clear")})
    print("   solve:", {k: "%.1f%%" % (100*p[k]/tot) for k in p if k.startswith("s_") and p[k]})
    print("   solve cycles/world-step:", {k: int(p[k]/ws) for k in p if k.startswith("s_") and p[k]}, flush=True)

if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [16384]
    for n in sizes:
        run(n, "f32", int(os.environ.get("SWEEP_LANES", "4")))
    if os.environ.get("PROFILE_OCC"):
        for ctas in (1, 5, 10):
            run(148 * ctas * 8, "f32", 4, settle=300, ctas=ctas)
