"""Small run for compute-sanitizer: 24 jittered pyramid worlds (fp32) through free fall and landing, a mixed scene, and deep stacks
(balanced islands: joint island schedule, group-capable position kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for name, cfg, vi, pi in (("pyramid", 3, 8, 3), ("mixed", 4, 8, 3), ("deep_stacks", 5, 8, 3)):
    bw = b2.BatchedWorld.replicate(scenes.SCENES[name](b2), 24, jitter=1e-2, config=cfg, real="f32")
    bw.step(1 / 60, vi, pi, steps)
    torch.cuda.synchronize()
    print(name, "status", bw.status(False), "contacts", int(bw.t["w_count"][:, 2].max()))
