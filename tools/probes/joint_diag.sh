# development probe: per-phase cycles of a scene with two builds of the kernels (variants/libb2b_<name>.so)
for v in head new; do
python - "$v" <<'PY'
import sys, os, runpy
sys.path.insert(0, "/root/repo")
from pypybox2d_b200 import _lib
_lib.SO_PATH = os.path.join(_lib.CSRC, "variants", "libb2b_%s.so" % sys.argv[1])
print("== variant", sys.argv[1])
for a in (["pyramid", "1", "23680"],):
    sys.argv = ["x"] + a
    runpy.run_path("/root/repo/tools/gpu_scene_profile.py", run_name="__main__")
PY
done
