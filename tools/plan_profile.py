"""Development probe: where the plan kernel's cycles go on the scenes that replan every step (per-phase cycle counters of thread 0:
collide ordering / destruction, adjacency, island walk, level schedule)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema
import bench

T = bench._scene_table()
for name in sys.argv[1:] or ["vertical_stack", "deep_stacks", "mixed"]:
    S = T[name]
    n = min(S["worlds"], 16384)
    w = scenes.SCENES[name](b2)
    bw = b2.BatchedWorld.replicate(w, n, jitter=1e-2, config=S["config"], real="f32", caps=S["caps"])
    bw.step(S["dt"], S["vi"], S["pi"], 300)
    torch.cuda.synchronize()
    bw.step(S["dt"], S["vi"], S["pi"], 8, profile=True)
    torch.cuda.synchronize()
    p = bw.profile()
    ws = max(p.pop("world_steps"), 1)
    keys = ("pairs0", "collide", "adjacency", "islands", "s_schedule", "sync", "pairs")
    print(name, "worlds", n, "cycles per world-step of thread 0:", {k: int(p[k] / ws) for k in keys}, "status", bw.status(False), flush=True)
