import sys
import os; sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np, torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes
# all nine joint types, fp32 and fp64, a few worlds, 40 steps from the scripted start (landing, limits reached, sleepers woken)
for real in ("f32", "f64"):
    ws = [scenes.joints(b2, world=k) for k in range(3)] + [scenes.bridge(b2, world=1), scenes.hello(b2)]
    bw = b2.BatchedWorld.from_worlds(ws, real=real)
    bw.step(1/60, 8, 3, 40)
    torch.cuda.synchronize()
    print(real, "status", bw.status(False), "types", bin(bw._caps_struct.joint_types), "finite", bool(np.isfinite(bw.download(["b_pos"])["b_pos"]).all()))
w = scenes.bridge(b2)
bw = b2.BatchedWorld.replicate(w, 64, jitter=1e-2, config=10, real="f32")
bw.step(1/60, 8, 3, 40)
torch.cuda.synchronize()
print("bridge revolute-only", bw.status(False), bin(bw._caps_struct.joint_types))
