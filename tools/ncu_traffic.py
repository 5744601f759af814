"""DRAM traffic of the step kernels for bench.py's `roofline.traffic`: reads one `ncu --set full` (or --metrics dram__bytes*)
CSV of a bench.py run and writes profiles/r2_traffic.json keyed by scene/worlds/kernel.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:b2b:: \
      --csv --log-file gpurun_out/traffic.csv python bench.py --worlds N --steps 1 --warmup 3 --no-cpu-baseline --t0-worlds 0 --no-invariance
  python tools/ncu_traffic.py gpurun_out/traffic.csv pyramid N

The LAST launch of every kernel (a settled step) is the one recorded."""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, scene, worlds = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
head = rows[0]
ki, mi, vi, ui = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value"), head.index("Metric Unit")
idi = head.index("ID")
last = {}
for r in rows[1:]:
    m = re.search(r"(collide|plan|prepare|solve_velocity|solve_position|sync|broadphase|toi)(?:_joints)?_kernel_", r[ki])
    if not m:
        continue
    k = "pairs" if m.group(1) == "broadphase" else m.group(1)
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
    last.setdefault(k, {}).setdefault(int(r[idi]), {})[r[mi]] = v * scale
out_path = os.path.join(ROOT, "profiles", "r2_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for k, launches in last.items():
    lid = max(launches)
    d = launches[lid]
    out["%s/%d/%s" % (scene, worlds, k)] = {
        "dram_bytes_per_launch": d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0),
        "dram_read": d.get("dram__bytes_read.sum"), "dram_write": d.get("dram__bytes_write.sum"),
        "duration_ns_under_ncu": d.get("gpu__time_duration.sum"),
        "how": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the last launch of the kernel in `bench.py --scene %s --worlds %d` "
               "(tools/ncu_traffic.py)" % (scene, worlds)}
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in out.items()}, indent=1))
