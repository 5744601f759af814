"""Development probe: where do the cycles of a velocity level-step go?  Needs a build with -DB2B_TRACE
(tools/build_variant.sh trace -DB2B_TRACE=1).  Segments, lane 0 of every team, summed over the run:
 0 wait for the current record | 1 table + issue of the next record's loads | 2 wait for the bodies' state (LDS)
 3 arithmetic of the constraint | 4 stores (+ the whole warm-start solve) | 5 barrier | 6 stamp overhead"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypybox2d_b200 import _lib
_lib.SO_PATH = os.path.join(_lib.CSRC, "variants", "libb2b_trace.so")
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema
for n, ctas in ((11840, 0), (1184, 1)):
    bw = b2.BatchedWorld.replicate(scenes.pyramid(b2), n, jitter=1e-2, config=3, real="f32", caps=schema.Caps(211, 211, 640, npair=1280),
                                   threads=4, lanes_aux=8, ctas_per_sm=ctas)
    bw.step(1 / 60, 8, 3, 300)
    torch.cuda.synchronize()
    out = (ctypes.c_int64 * 32)()
    _lib.check(_lib.lib().b2b_profile(bw._ctx, out, 1), "b2b_profile")
    steps = 4
    bw.step(1 / 60, 8, 3, steps)
    _lib.check(_lib.lib().b2b_profile(bw._ctx, out, 1), "b2b_profile")
    seg = [out[20 + i] for i in range(7)]
    level_steps = n * steps * 9 * 274.0
    print("worlds %d (ctas/SM %s): cycles per level-step and world:" % (n, ctas or "max"),
          {k: round(v / level_steps, 1) for k, v in zip(("rec_wait", "issue_next", "lds_wait", "fp", "stores", "barrier", "stamp"), seg)},
          "sum %.0f" % (sum(seg) / level_steps), flush=True)
