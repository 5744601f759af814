"""Development probe: where do two lane-team sizes diverge? (T4 invariance debugging)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes
w = scenes.pyramid(b2)
la, lb = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 32
A = b2.BatchedWorld.replicate(w, 16, jitter=1e-2, config=3, real="f32", threads=la)
B = b2.BatchedWorld.replicate(w, 16, jitter=1e-2, config=3, real="f32", threads=lb)
for s in range(90):
    A.step(1 / 60, 8, 3, 1)
    B.step(1 / 60, 8, 3, 1)
    a, b = A.download(["b_pos", "b_vel", "c_imp", "w_count"]), B.download(["b_pos", "b_vel", "c_imp", "w_count"])
    d = {k: float(np.nanmax(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)))) for k in a}
    if any(v != 0 for v in d.values()):
        print("step", s, d, "contacts", a["w_count"][0]); 
        if s > 40: break
print("done")
