"""Development probe #2 (GPU box): fp32 teacher-forced / settled parity numbers and steady-state timing."""
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

INT_FIELDS = ["w_count", "f_moved", "c_fix", "c_flags", "c_mtype", "b_flags"]
REAL_FIELDS = [f[0] for f in schema.FIELDS if f[3] == "real" and f[1] != "S"]


def round32(A):
    B = {k: v.copy() for k, v in A.items()}
    for k in REAL_FIELDS:
        B[k] = B[k].astype(np.float32).astype(np.float64)
    return B


def ints_equal(OA, G):
    bad = harness.diff_state(OA, G, fields=INT_FIELDS)
    if harness.oracle_islands(OA) != harness.oracle_islands(G):
        bad.append("islands")
    if harness.oracle_pairs(OA) != harness.oracle_pairs(G):
        bad.append("pairs")
    # manifold ids of used points
    nc = int(OA["w_count"][0][2])
    cnt = (OA["c_mtype"][0][:nc] >> 8) & 0xff
    for j in range(2):
        m = cnt > j
        if not np.array_equal(OA["c_id"][0][:nc][m, j], G["c_id"][0][:nc][m, j]):
            bad.append("ids")
    return bad


def teacher_forced(name, wi, steps, dt, vi, pi, stride=1):
    mw, OA, caps = harness.facade_arrays(name, wi)
    orc = oracle.Oracle(OA, 0)
    bw = b2.BatchedWorld({k: v.copy() for k, v in OA.items()}, caps, real="f32")
    n_bad = 0
    worst = {"b_pos": 0.0, "b_vel": 0.0, "c_imp": 0.0}
    first_bad = None
    inv = 0.0
    for s in range(steps):
        if s % stride == 0:
            T = round32(OA)
            o2 = oracle.Oracle(T, 0)
            o2.inv_dt0 = inv
            bw.upload({k: T[k] for k in schema.FIELD_NAMES if not k.startswith("s_")})
            bw.inv_dt0 = inv
            o2.step(dt, vi, pi)
            bw.step(dt, vi, pi)
            G = bw.download()
            bad = ints_equal(T, G)
            if bad:
                n_bad += 1
                if first_bad is None:
                    first_bad = (s, bad)
            else:
                for k in worst:
                    x, y = harness.used_view(T, k), harness.used_view(G, k).astype(np.float64)
                    if k == "c_imp":
                        nc = x.shape[0]
                        cnt = (T["c_mtype"][0][:nc] >> 8) & 0xff
                        touching = (T["c_flags"][0][:nc] & 1) != 0
                        mask = np.zeros_like(x, dtype=bool)
                        mask[:, 0:2] = ((cnt > 0) & touching)[:, None]
                        mask[:, 2:4] = ((cnt > 1) & touching)[:, None]
                        d = np.abs(x - y)[mask]
                    else:
                        d = np.abs(x[:, :3] - y[:, :3])
                    if d.size:
                        worst[k] = max(worst[k], float(d.max()))
        orc.step(dt, vi, pi)
        inv = 1.0 / dt
    print("teacher-forced %-15s w%d f32: %d/%d steps with integer mismatch (first %s); worst abs diffs %s" % (
        name, wi, n_bad, steps // stride, first_bad, {k: "%.2e" % v for k, v in worst.items()}), flush=True)


def settled_free_run(name, wi, settle, steps, dt, vi, pi, real):
    mw, OA, caps = harness.facade_arrays(name, wi)
    orc = oracle.Oracle(OA, 0)
    orc.step(dt, vi, pi, settle)
    T = round32(OA) if real == "f32" else {k: v.copy() for k, v in OA.items()}
    o2 = oracle.Oracle(T, 0)
    o2.inv_dt0 = 1.0 / dt
    bw = b2.BatchedWorld({k: v.copy() for k, v in T.items()}, caps, real=real)
    bw.inv_dt0 = 1.0 / dt
    first_bad = None
    worst = 0.0
    for s in range(steps):
        o2.step(dt, vi, pi)
        bw.step(dt, vi, pi)
        G = bw.download()
        bad = ints_equal(T, G)
        if bad and first_bad is None:
            first_bad = (s, bad)
            break
        for k in ("b_pos", "b_vel"):
            x, y = harness.used_view(T, k), harness.used_view(G, k).astype(np.float64)
            worst = max(worst, float(np.abs(x[:, :3] - y[:, :3]).max()))
    print("settled(%d)+%d %-15s w%d %s: first integer mismatch %s; worst abs state diff %.2e" % (
        settle, steps, name, wi, real, first_bad, worst), flush=True)


def timing_settled(n_worlds, real, settle=240, steps=10, threads=0, name="pyramid", config=3):
    w = scenes.SCENES[name](b2)
    nd = sum(1 for b in w.bodies if b.type == 1)
    bw = b2.BatchedWorld.replicate(w, n_worlds, jitter=1e-2, config=config, real=real, threads=threads)
    torch.cuda.synchronize()
    t0 = time.time()
    bw.step(1 / 60, 8, 3, settle)
    torch.cuda.synchronize()
    t_settle = time.time() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bw.step(1 / 60, 8, 3, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    D = bw.download(["w_count", "w_island"], worlds=slice(0, 64))
    print("timing settled %s %s worlds=%d threads=%d: setup %d steps %.1f s (%.1f ms/step); timed %.2f ms/step, "
          "%.3e body-steps/s; grid=%s nc~%.0f islands~%.1f status=%d" % (
              name, real, n_worlds, threads, settle, t_settle, 1e3 * t_settle / settle, ms / steps,
              nd * n_worlds * steps / (ms * 1e-3), bw.info()["grid"], D["w_count"][:, 2].mean(), D["w_island"][:, 0].mean(),
              bw.status(False)), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["tf", "settled", "timing"]
    if "timing" in which:
        for th in (32, 64, 128):
            timing_settled(4096, "f32", threads=th)
        timing_settled(4096, "f64", threads=64)
        timing_settled(16384, "f32", threads=64, settle=240, steps=5)
    if "tf" in which:
        teacher_forced("pyramid", 1, 300, 1 / 60, 8, 3)
        teacher_forced("pyramid", 0, 150, 1 / 60, 8, 3)
        teacher_forced("mixed", 1, 150, 1 / 60, 8, 3)
        teacher_forced("vertical_stack", 1, 150, 1 / 60, 10, 10)
    if "settled" in which:
        settled_free_run("pyramid", 1, 300, 300, 1 / 60, 8, 3, "f32")
        settled_free_run("pyramid", 0, 300, 300, 1 / 60, 8, 3, "f32")
        settled_free_run("pyramid", 1, 0, 600, 1 / 60, 8, 3, "f64")
        settled_free_run("mixed", 1, 0, 400, 1 / 60, 8, 3, "f64")
        settled_free_run("deep_stacks", 1, 0, 400, 1 / 60, 8, 3, "f64")
        settled_free_run("vertical_stack", 1, 0, 400, 1 / 60, 10, 10, "f64")
