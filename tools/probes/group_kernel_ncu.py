"""Development probe: the two position kernels on the same settled pyramid batch (the group-capable one forced with
B2B_STEP_GROUP_KERNEL), for `ncu -k regex:solve_position --launch-skip 600 --launch-count 8`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema
bw = b2.BatchedWorld.replicate(scenes.pyramid(b2), 11840, jitter=1e-2, config=3, real="f32", caps=schema.Caps(211, 211, 640, npair=1280),
                               threads=4, lanes_aux=8)
bw.step(1 / 60, 8, 3, 300)
bw.step(1 / 60, 8, 3, 2)
bw.step(1 / 60, 8, 3, 2, group_kernel=True)
torch.cuda.synchronize()
print("status", bw.status(False))
