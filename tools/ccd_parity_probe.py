"""Measure the continuous-collision (TOI) parity tiers on the device for every CCD golden case (development aid).

free: fp64 kernels free-running from t = 0 against the oracle: first step with an integer mismatch, worst float difference
up to it.  one64 / one32: teacher-forced ONE step from the oracle's state at every step k: steps with integer mismatches,
worst float differences on the others.  Writes gpurun_out/ccd_parity.json.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ.setdefault("B2B_PARITY_REPORT_ONLY", "1")
import oracle  # noqa: E402
from conftest import golden_cases, load_golden, arrays_from_golden  # noqa: E402
import test_gpu_parity as T  # noqa: E402
from pypybox2d_b200 import schema  # noqa: E402
import pypybox2d_b200 as b2  # noqa: E402


def main():
    out = {}
    cases = [c for c in golden_cases() if "ccd" in c or "bullets" in c]
    cases = [c for c in cases if not sys.argv[1:] or c in sys.argv[1:]]
    names = [k for k in schema.FIELD_NAMES if not k.startswith("s_")]
    for case in cases:
        g = load_golden(case)
        dt, vi, pi, steps, world = g["meta"]
        vi, pi, steps = int(vi), int(pi), int(steps)
        rec = {}
        # free-running fp64
        A, caps = arrays_from_golden(g)
        bw = b2.BatchedWorld({k: v.copy() for k, v in A.items()}, caps, real="f64")
        orc = oracle.Oracle(A, 0)
        worst, first_bad = 0.0, None
        for s in range(steps):
            orc.step(dt, vi, pi)
            bw.step(dt, vi, pi)
            G = bw.download()
            bad = T.int_mismatches(A, G)
            if bad:
                first_bad = [s, bad]
                break
            worst = max(worst, max(T.float_diffs(A, G).values()))
        rec["free64"] = {"first_int_mismatch": first_bad, "worst_float": worst, "steps": steps, "status": bw.status(False)}
        # teacher-forced one step, fp64 and fp32
        for real in ("f64", "f32"):
            A, caps = arrays_from_golden(g)
            orc = oracle.Oracle(A, 0)
            bw = b2.BatchedWorld({k: v.copy() for k, v in A.items()}, caps, real=real)
            inv, worst, ties, worst_p, worst_v = 0.0, 0.0, [], 0.0, 0.0
            for s in range(steps):
                S = A if real == "f64" else T.round32(A)
                bw.upload({k: S[k] for k in names})
                bw.inv_dt0 = inv
                if real == "f32":
                    o2 = oracle.Oracle(S, 0)
                    o2.inv_dt0 = inv
                    o2.step(dt, vi, pi)
                orc.step(dt, vi, pi)
                bw.step(dt, vi, pi)
                G = bw.download()
                R = A if real == "f64" else S
                bad = T.int_mismatches(R, G)
                if bad:
                    ties.append([s, bad])
                else:
                    if real == "f64":
                        worst = max(worst, max(T.float_diffs(R, G).values()))
                    else:
                        dp, dv = T.state_diff(R, G)
                        worst_p, worst_v = max(worst_p, dp), max(worst_v, dv)
                inv = 1.0 / dt
            rec["one_" + real] = {"tie_steps": len(ties), "ties": ties[:12], "worst_float": worst, "worst_p": worst_p, "worst_v": worst_v,
                                  "steps": steps, "status": bw.status(False)}
        out[case] = rec
        print(case, json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ccd_parity.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
