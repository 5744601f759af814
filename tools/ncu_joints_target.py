"""Settled bridge batch for ncu (joints in the island solver): settle N worlds (untimed launches), then ONE profiled single-step launch.

  ncu --set full --clock-control none --import-source on -k regex:solve_velocity_joints --launch-skip 1 --launch-count 1 \
      -o gpurun_out/joints_velocity python tools/ncu_joints_target.py 131072
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
real = sys.argv[2] if len(sys.argv) > 2 else "f32"
w = scenes.bridge(b2)
caps = schema.Caps(39, 39, 220, nj=31)
seed = b2.BatchedWorld.replicate(w, 2048, jitter=1e-2, config=10, real=real, caps=caps)
seed.step(1 / 60, 8, 3, 300)      # one launch of every kernel per step; the velocity kernel's first profiled launch is skipped below
torch.cuda.synchronize()
bw = seed.tiled(n)
bw.step(1 / 60, 8, 3, 1)
torch.cuda.synchronize()
bw.step(1 / 60, 8, 3, 1)          # the profiled one
torch.cuda.synchronize()
print("done", bw.info(), bw.status(False))
