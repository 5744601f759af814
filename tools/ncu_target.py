"""Small settled-pyramid run for ncu: settle N worlds (untimed launches), then ONE profiled single-step launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
real = sys.argv[2] if len(sys.argv) > 2 else "f32"
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 32
w = scenes.pyramid(b2)
caps = schema.Caps(211, 211, 640, npair=1280)
bw = b2.BatchedWorld.replicate(w, n, jitter=1e-2, config=3, real=real, caps=caps, threads=threads)
bw.step(1/60, 8, 3, 200)      # launch 2 (launch 1 is init_proxies): settle
torch.cuda.synchronize()
bw.step(1/60, 8, 3, 1)        # launch 3: the profiled one
torch.cuda.synchronize()
print("done", bw.info(), bw.status(False))
