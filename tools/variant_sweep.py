"""Development aid: time the settled-pyramid step with experimental builds of the kernels (tools/build_variant.sh).
usage: variant_sweep.py <n_worlds> <name> [<name> ...]   ('base' = the committed libb2b.so); one subprocess per variant."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    from pypybox2d_b200 import _lib
    name = sys.argv[3]
    if name != "base":
        _lib.SO_PATH = os.path.join(_lib.CSRC, "variants", "libb2b_%s.so" % name)
    import gpu_sweep
    print("variant", name, end=": ", flush=True)
    gpu_sweep.run(int(sys.argv[2]), "f32", int(os.environ.get("SWEEP_LANES", "4")), 8, ctas=int(os.environ.get("SWEEP_CTAS", "0")))
else:
    for name in sys.argv[2:]:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--one", sys.argv[1], name], cwd=os.path.join(ROOT, "tools"))
