"""Pyramid runs for ncu captures of the TOI kernel: `settled` (settle with continuous collision off, switch it on, ONE
profiled step: every world looks at its 20 ground contacts and finds no event) or `landing` (continuous collision on from
the scripted start, `skip` untimed steps, then ONE profiled step in the landing transient: events in most worlds)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema
mode = sys.argv[1] if len(sys.argv) > 1 else "settled"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 45
caps = schema.Caps(211, 211, 640, npair=1280)
if mode == "settled":
    bw = b2.BatchedWorld.replicate(scenes.pyramid(b2), n, jitter=1e-2, config=3, real="f32", caps=caps)
    bw.step(1 / 60, 8, 3, 200)
    bw.set_continuous(True)
    bw.step(1 / 60, 8, 3, 2)
else:
    bw = b2.BatchedWorld.replicate(scenes.pyramid(b2, continuous_physics=True), n, jitter=1e-2, config=3, real="f32", caps=caps)
    bw.step(1 / 60, 8, 3, skip)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
bw.step(1 / 60, 8, 3, 1, time_kernels=True)        # the profiled step
e1.record()
torch.cuda.synchronize()
print("done", mode, n, "step ms", e0.elapsed_time(e1), bw.kernel_ms(), bw.status(False))
