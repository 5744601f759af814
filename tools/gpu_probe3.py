"""Development probe #3: anatomy of fp32 integer mismatches from the settled pyramid."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import harness, oracle, torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import schema
from gpu_probe2 import round32, ints_equal

def run(name, wi, settle, steps, dt, vi, pi):
    mw, OA, caps = harness.facade_arrays(name, wi)
    orc = oracle.Oracle(OA, 0); orc.step(dt, vi, pi, settle)
    T = round32(OA)
    o2 = oracle.Oracle(T, 0); o2.inv_dt0 = 1/dt
    bw = b2.BatchedWorld({k: v.copy() for k, v in T.items()}, caps, real="f32"); bw.inv_dt0 = 1/dt
    nbad = 0
    for s in range(steps):
        o2.step(dt, vi, pi); bw.step(dt, vi, pi)
        G = bw.download()
        bad = ints_equal(T, G)
        x, y = harness.used_view(T, "b_pos"), harness.used_view(G, "b_pos").astype(np.float64)
        dpos = np.abs(x[:, :3]-y[:, :3]).max()
        if bad:
            nbad += 1
            nc = int(T["w_count"][0][2])
            mt_o, mt_g = T["c_mtype"][0][:nc], G["c_mtype"][0][:nc]
            idx = np.nonzero(mt_o != mt_g)[0]
            if nbad <= 6:
                print("step", s, bad, "dpos %.2e" % dpos, "contacts differing:", idx[:6].tolist())
                for c in idx[:3]:
                    print("   c%d fix=%s oracle type/count=%d/%d gpu=%d/%d ids o=%s g=%s flags o=%x g=%x" % (
                        c, T["c_fix"][0][c].tolist(), mt_o[c] & 255, mt_o[c] >> 8, mt_g[c] & 255, mt_g[c] >> 8,
                        [hex(v) for v in T["c_id"][0][c]], [hex(v) for v in G["c_id"][0][c]], T["c_flags"][0][c], G["c_flags"][0][c]))
                    print("      oracle mf", T["c_mf"][0][c], "pts", T["c_pts"][0][c])
                    print("      gpu    mf", G["c_mf"][0][c], "pts", G["c_pts"][0][c])
    x, y = harness.used_view(T, "b_pos"), harness.used_view(G, "b_pos").astype(np.float64)
    print("%s w%d: %d/%d steps with integer mismatch; final max |dpos| %.2e, |dvel| %.2e; touching o=%d g=%d" % (
        name, wi, nbad, steps, np.abs(x[:, :3]-y[:, :3]).max(),
        np.abs(harness.used_view(T, "b_vel")[:, :3]-harness.used_view(G, "b_vel")[:, :3]).max(),
        (harness.used_view(T,"c_flags") & 1).sum(), (harness.used_view(G,"c_flags") & 1).sum()))

run("pyramid", 1, 300, 300, 1/60, 8, 3)
run("pyramid", 0, 300, 300, 1/60, 8, 3)
