"""Where do the fp64 kernels' free-running differences from the oracle come from (5e-14 on the pyramid,
8.7e-7 on the mixed scene)?  CPU experiment with the host-compiled device source (tests/hostcheck):
  (1) built as is (glibc sin / cos, libm pow for the squares the reference writes `x ** 2`) it reproduces the oracle bit
      for bit on every scene (that is tests/test_hostcheck.py);
  (2) -DB2B_SINCOS_ULP_PROBE: about half of the sin / cos results moved by ONE ulp, nothing else changed (what CUDA's
      sincos, <= 2 ulp from glibc's, does);
  (3) -DB2B_SQ_PRODUCT: the correctly rounded product x * x where the reference calls pow(x, 2.0) (what the device does:
      glibc's pow differs from the product by one ulp in 0.09 % of evaluations).
  The free-running differences of (2) and (3) after the golden trajectory length are the sizes the GPU's fp64 kernels
  show against the oracle: ~1e-13 where the pile is quiet, up to ~1e-6 where bodies roll and tumble.
usage: python tools/sincos_ulp_probe.py"""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "hostcheck"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import harness, oracle, hostcheck

def build(flag):
    so = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck_probe%d.so" % len(flag + [""] * (1 if flag and "SQ" in flag[0] else 0)))
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-builtin-pow"] + flag +
                          ["-o", so, os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp"), "-lm"])
    L = ctypes.CDLL(so)
    L.hc_step.argtypes = [ctypes.POINTER(hostcheck.Caps), ctypes.POINTER(hostcheck.Buffers), ctypes.c_double, ctypes.c_double,
                          ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return L

def run(L, scene, world, dt, vi, pi, steps):
    mw, OA, caps = harness.facade_arrays(scene, world)
    HA = {k: v.copy() for k, v in OA.items()}
    orc = oracle.Oracle(OA, 0)
    hostcheck._LIB = L
    inv, worst, first_int = 0.0, 0.0, None
    for s in range(steps):
        orc.step(dt, vi, pi)
        hostcheck.step(HA, dt, inv, vi, pi)
        inv = 1.0 / dt
        if first_int is None and harness.diff_state(OA, HA, fields=["w_count", "c_fix", "c_flags", "c_mtype"]):
            first_int = s
        if first_int is None:
            d = max(np.abs(harness.used_view(OA, k)[:, :3] - harness.used_view(HA, k)[:, :3]).max() for k in ("b_pos", "b_vel"))
            worst = max(worst, float(d))
    return worst, first_int

for name, flag in (("as the oracle", []), ("sin/cos +-1 ulp", ["-DB2B_SINCOS_ULP_PROBE=1"]), ("x*x for x**2", ["-DB2B_SQ_PRODUCT=1"])):
    L = build(flag)
    for scene, world, dt, vi, pi, steps in (("pyramid", 1, 1 / 60, 8, 3, 600), ("mixed", 1, 1 / 60, 8, 3, 400),
                                             ("vertical_stack", 1, 1 / 60, 10, 10, 400), ("deep_stacks", 1, 1 / 60, 8, 3, 400)):
        w, fi = run(L, scene, world, dt, vi, pi, steps)
        print("%-14s %-15s %d steps: max |d(pos, vel)| vs oracle %.2e%s" % (name, scene, steps, w,
              "" if fi is None else " (integer outputs differ from step %d on)" % fi), flush=True)
