"""Development probe: per-phase cycles of the plan-side kernels for a scene in its 'settled' regime."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes
name, cfg, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
vi, pi = (10, 10) if name in ("vertical_stack", "hello") else (8, 3)
bw = b2.BatchedWorld.replicate(scenes.SCENES[name](b2), n, jitter=1e-2, config=cfg, real="f32", lanes_aux=int(os.environ.get("LANES_AUX", "0")))
bw.step(1 / 60, vi, pi, 300)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); bw.step(1 / 60, vi, pi, 10, time_kernels=True); e1.record(); torch.cuda.synchronize()
k = bw.kernel_ms()
print(name, "ms/step %.2f" % (e0.elapsed_time(e1) / 10), {q: round(k[q] / 10, 2) for q in ("collide", "plan", "solve", "broadphase")}, bw.info(), bw.caps.__dict__ if hasattr(bw.caps, "__dict__") else bw.caps)
bw.step(1 / 60, vi, pi, 10, profile=True)
p = bw.profile(); ws = max(p.pop("world_steps"), 1)
print("cycles/world-step:", {q: int(v / ws) for q, v in p.items() if v}, "world_steps", ws)
D = bw.download(["w_island", "w_count"], worlds=slice(0, 256))
print("islands per world %.1f, island bodies %.1f, island contacts %.1f, contacts %.1f" % (D["w_island"][:, 0].mean(), D["w_island"][:, 1].mean(), D["w_island"][:, 2].mean(), D["w_count"][:, 2].mean()))
