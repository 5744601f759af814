"""Development probe (GPU box): CUDA path vs the C oracle on a few scenes + a first timing."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import harness  # noqa
import oracle  # noqa
import torch  # noqa
import pypybox2d_b200 as b2  # noqa
from pypybox2d_b200 import schema, scenes  # noqa

INT_FIELDS = [k for k in harness.STATE_FIELDS if k.startswith(("w_", "f_moved", "c_fix", "c_flags", "c_mtype", "b_flags"))]


def compare(name, wi, steps, dt, vi, pi, real, tol):
    mw, OA, caps = harness.facade_arrays(name, wi)
    bw = b2.BatchedWorld({k: v.copy() for k, v in OA.items()}, caps, real=real)
    orc = oracle.Oracle(OA, 0)
    first = None
    worst = 0.0
    for s in range(steps):
        orc.step(dt, vi, pi)
        bw.step(dt, vi, pi)
        G = bw.download()
        bad = harness.diff_state(OA, G, fields=INT_FIELDS)
        if harness.oracle_islands(OA) != harness.oracle_islands(G):
            bad.append("islands")
        if harness.oracle_pairs(OA) != harness.oracle_pairs(G):
            bad.append("pairs")
        for k in ("b_pos", "b_vel", "b_xf"):
            x, y = harness.used_view(OA, k), harness.used_view(G, k).astype(np.float64)
            d = float(np.abs(x - y).max())
            worst = max(worst, d)
            if d > tol:
                bad.append("%s:%.2e" % (k, d))
        if bad and first is None:
            first = (s, bad)
            break
    print("%-16s w%d %s steps=%d first_mismatch=%s worst=%.2e nc=%d/%d status=%d" % (
        name, wi, real, s + 1, first, worst, OA["w_count"][0][2], G["w_count"][0][2], bw.status(False)), flush=True)


def timing(n_worlds, real, steps=20, threads=0):
    w = scenes.pyramid(b2)
    bw = b2.BatchedWorld.replicate(w, n_worlds, jitter=1e-2, config=3, real=real, threads=threads)
    torch.cuda.synchronize()
    bw.step(1 / 60, 8, 3, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bw.step(1 / 60, 8, 3, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("timing %s worlds=%d threads=%d steps=%d: %.2f ms/step, %.3e body-steps/s info=%s status=%d" % (
        real, n_worlds, threads, steps, ms / steps, 210.0 * n_worlds * steps / (ms * 1e-3), bw.info(), bw.status(False)), flush=True)
    return bw


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    compare("hello", 0, 60, 0.1, 10, 10, "f64", 1e-9)
    compare("pyramid", 1, 120, 1 / 60, 8, 3, "f64", 1e-8)
    compare("pyramid", 0, 120, 1 / 60, 8, 3, "f64", 1e-8)
    compare("vertical_stack", 1, 120, 1 / 60, 10, 10, "f64", 1e-8)
    compare("mixed", 1, 120, 1 / 60, 8, 3, "f64", 1e-8)
    compare("deep_stacks", 1, 120, 1 / 60, 8, 3, "f64", 1e-8)
    compare("pyramid", 1, 60, 1 / 60, 8, 3, "f32", 1e-3)
    for th in (32, 64, 128):
        timing(4096, "f32", 20, th)
    timing(4096, "f64", 10, 64)
