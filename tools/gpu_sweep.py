"""Development probe: throughput of settled pyramid worlds for lane-team configurations, with per-kernel ms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema

def run(n_worlds, real, lanes, lanes_aux, settle=300, steps=5, repeats=1, ctas=0):
    w = scenes.pyramid(b2)
    caps = schema.Caps(211, 211, 640, npair=1280)
    bw = b2.BatchedWorld.replicate(w, n_worlds, jitter=1e-2, config=3, real=real, caps=caps, threads=lanes, lanes_aux=lanes_aux, ctas_per_sm=ctas)
    bw.step(1/60, 8, 3, settle)
    torch.cuda.synchronize()
    for pf in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bw.step(1/60, 8, 3, steps, time_kernels=True, group_kernel=bool(int(os.environ.get('SWEEP_GROUP_KERNEL', '0')))); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        k = bw.kernel_ms()
        print("%s worlds=%d lanes=%d aux=%d ctas=%d: %.2f ms/step %.3e body-steps/s; kernels ms/step collide %.2f plan %.2f prepare %.2f velocity %.2f position %.2f broad %.2f status=%d" % (
            real, n_worlds, lanes, lanes_aux, ctas, ms/steps, 210*n_worlds*steps/(ms*1e-3), k["collide"]/steps, k["plan"]/steps, k["prepare"]/steps,
            k["solve_velocity"]/steps, k["solve_position"]/steps, k["broadphase"]/steps, bw.status(False)), flush=True)

if __name__ == "__main__":
    import sys
    sizes = [int(a) for a in sys.argv[1:]] or [32768]
    lanes_list = [int(x) for x in os.environ.get("SWEEP_LANES", "4").split(",")]
    for n in sizes:
        for lanes in lanes_list:
            run(n, "f32", lanes, 8)
