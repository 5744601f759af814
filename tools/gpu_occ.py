"""Development probe: solve-kernel time vs resident CTAs per SM (is a level-step latency bound or contended?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema

def run(n_worlds, ctas, lanes=4, settle=300, steps=5):
    w = scenes.pyramid(b2)
    caps = schema.Caps(211, 211, 640, npair=1280)
    bw = b2.BatchedWorld.replicate(w, n_worlds, jitter=1e-2, config=3, real="f32", caps=caps, threads=lanes, lanes_aux=8, ctas_per_sm=ctas)
    bw.step(1/60, 8, 3, settle)
    torch.cuda.synchronize()
    bw.step(1/60, 8, 3, steps, time_kernels=True); torch.cuda.synchronize()
    k = bw.kernel_ms()
    print("worlds=%d ctas/sm=%d lanes=%d: solve %.3f ms/step (%.1f ns/world), info %s" % (n_worlds, ctas, lanes, k["solve"]/steps, 1e6*k["solve"]/steps/n_worlds, bw.info()), flush=True)

if __name__ == "__main__":
    for ctas in (10, 8, 5, 2, 1):
        run(148 * ctas * 8, ctas)
