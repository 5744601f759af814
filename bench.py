#!/usr/bin/env python
"""bench.py -- body-steps/s of the batched World.step on 20-row pyramid worlds (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path, on the host cores
    python bench.py --scene {vertical_stack,mixed,deep_stacks,hello}   # the other BASELINE.json configurations
    python bench.py --scene {bridge,joints}                            # joints in the island solver

Workload (config.workload): the pyramid-20 scene of examples/pyramid-bench.py (210 dynamic boxes on an edge
ground), step(1/60, 8, 3), CCD off, `--worlds` independent worlds PER GPU (131,072 by default, i.e. 1,048,576
worlds at 8 GPUs: BASELINE configs[2], weak scaling, no collective on the step path).  Every world has its own
jittered initial positions.  The timed region runs in the SETTLED regime (one 211-body island, ~590 contacts,
~400 touching, the most expensive phase of the trajectory): `--distinct` distinct worlds are stepped `--settle`
steps untimed and that state is tiled over the batch.

A "step" is one World.step of every world.  `value` times K consecutive steps with state resident in HBM (CUDA events
on the launching stream, one per step so that min / median / max are visible; max over ranks); `e2e` times K calls of the
host-buffer C-ABI entry (b2b_step_host: pinned forces in, step, body state out, copies pipelined under the kernels).
`roofline` is the dominant kernel of THIS run (per-kernel CUDA events) against the measured HBM peak, with SURVEY §8d's
whole-step figure beside it.  `cpu_baseline` is the unmodified reference (oracle/_ref) on the host cores for a bounded
sample, with the C oracle port beside it -- a baseline, not the target.  `invariance` is the end-of-run T4 check: per-world
checksums of a sharded 64-world batch gathered over NCCL and compared with a one-GPU run (outside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT, VEL_ITERS, POS_ITERS = 1.0 / 60.0, 8, 3
N_DYNAMIC = 210
METRIC = "body-steps/sec (batched 20-row pyramid worlds)"


def _scene_table():
    from pypybox2d_b200 import schema
    C = schema.Caps
    return {   # BASELINE.json configs; worlds = the per-GPU size of the config
        "pyramid": dict(what="pyramid-20 (examples/pyramid-bench.py scene)", config=3, caps=C(211, 211, 640, npair=1280), dt=DT,
                        dt_text="1/60", vi=8, pi=3, n_dyn=210, worlds=131072, distinct=8192),
        "vertical_stack": dict(what="examples/verticalstack-bench.py scene (5 columns x 16 rows of boxes and circles)", config=2,
                               caps=C(161, 162, 512, npair=1424), dt=DT, dt_text="1/60", vi=10, pi=10, n_dyn=160, worlds=65536, distinct=2048),
        "hello": dict(what="examples/hello-bench.py scene (one circle on a box)", config=2, caps=None, dt=0.1, dt_text="0.1",
                      vi=10, pi=10, n_dyn=1, worlds=65536, distinct=1024),
        "mixed": dict(what="mixed circles + convex polygons, 256 bodies per world, per-world shape draws (SURVEY 8d config 4)",
                      config=4, caps=C(257, 260, 1104), dt=DT, dt_text="1/60", vi=8, pi=3, n_dyn=256, worlds=131072, distinct=512,
                      per_world_shapes=True),
        "deep_stacks": dict(what="10 columns x 20 boxes (SURVEY 8d config 5)", config=5, caps=C(201, 201, 1024, npair=2560), dt=DT, dt_text="1/60", vi=8,
                            pi=3, n_dyn=200, worlds=32768, distinct=2048),
        # not a BASELINE config: the joint hooks of the island solver, 30 planks on 31 revolute joints under 8 loose bodies
        "bridge": dict(what="plank bridge of 31 revolute joints under 8 loose bodies (testbed/bridge.py style; SURVEY 8f #3)", config=10,
                       caps=C(39, 39, 220, nj=31), dt=DT, dt_text="1/60", vi=8, pi=3, n_dyn=38, worlds=131072, distinct=2048),
        # the joint coverage scene: 44 joints of all nine built types (the generic joint-capable kernels, with the type dispatch)
        "joints": dict(what="joint coverage scene: 65 dynamic bodies, 44 joints of nine types (bridge, chain, motorised arms, lift, car, pulley ...)",
                       config=9, caps=C(66, 66, 328, nj=44), dt=DT, dt_text="1/60", vi=8, pi=3, n_dyn=65, worlds=65536, distinct=1024),
    }


SCENE_TABLE = None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="pyramid", choices=["pyramid", "vertical_stack", "mixed", "deep_stacks", "hello", "bridge", "joints"],
                    help="BASELINE.json configuration (default: the headline pyramid batch)")
    ap.add_argument("--worlds", type=int, default=0, help="worlds per GPU (0: the configuration's own size, 131,072 for the pyramid)")
    ap.add_argument("--distinct", type=int, default=0, help="distinct jittered worlds settled, then tiled (0: per scene)")
    ap.add_argument("--no-invariance", action="store_true", help="skip the end-of-run NCCL gather / T4 check")
    ap.add_argument("--settle", type=int, default=300)
    ap.add_argument("--real", default="f32", choices=["f32", "f64"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: same as --steps")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--t0-worlds", type=int, default=65536, help="worlds of the secondary from-t=0 run (0: skip it)")
    ap.add_argument("--t0-steps", type=int, default=600)
    ap.add_argument("--ccd-steps", type=int, default=5, help="steps of the secondary run with continuous collision on (0: skip it)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arms
def _settled_oracle_worlds(n, settle):
    """n oracle worlds (distinct jitter) advanced to the settled regime -- the same workload as the GPU arm."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import harness
    import oracle
    worlds = []
    for i in range(n):
        mw, A, caps = harness.facade_arrays("pyramid", 1 + i, config=3)
        o = oracle.Oracle(A, 0)
        o.step(DT, VEL_ITERS, POS_ITERS, settle)
        worlds.append(o)
    return worlds


def cpu_baseline_port(seconds, settle=300):
    """The C oracle on the host cores: one thread per core, one settled pyramid world each (ctypes releases the
    GIL inside orc_run).  Returns the cpu_baseline object."""
    from concurrent.futures import ThreadPoolExecutor
    cores = max(1, min(os.cpu_count() or 1, 32))
    with ThreadPoolExecutor(cores) as ex:
        worlds = list(ex.map(lambda i: _settled_oracle_worlds(1, settle)[0], range(cores)))
        chunk = 50
        ex.submit(lambda: None).result()
        t0 = time.perf_counter()
        done = 0
        while time.perf_counter() - t0 < seconds:
            list(ex.map(lambda o: o.step(DT, VEL_ITERS, POS_ITERS, chunk), worlds))
            done += chunk
        dt = time.perf_counter() - t0
    return {"value": N_DYNAMIC * cores * done / dt, "unit": "body-steps/s", "cores": cores, "kind": "port",
            "sample": "%d settled pyramid-20 worlds (one per thread) x %d steps of the C oracle (oracle/b2_oracle.c, "
                      "fp64, -O2), step(1/60,8,3), CCD off" % (cores, done)}


REF_WORKER = r"""
import sys, os, time, io, contextlib
sys.path.insert(0, %(ref)r)
core, settle, steps, warmup = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
pure, ccd = sys.argv[5] == "pure", sys.argv[6] == "ccd"
if pure:
    sys.modules["pypybox2d._common"] = None      # the reference's own fallback: common.py:954-979 "Using Pure Python"
with contextlib.redirect_stdout(io.StringIO()):
    import pypybox2d as b2
try:
    os.sched_setaffinity(0, {core})
except Exception:
    pass
from copy import copy
world = b2.World((0, -10), True, continuous_physics=ccd)
world.create_static_body(shapes=b2.Edge((-40, 0), (40, 0)))
fixture = b2.Fixture(b2.Polygon(box=(0.5, 0.5)), density=5.0)
x = b2.Vec2(-7, 0.75)
for i in range(20):
    y = copy(x)
    for j in range(i, 20):
        world.create_dynamic_body(position=y, fixtures=fixture)
        y += (1.125, 0)
    x += (0.5625, 1.25)
for i in range(settle + warmup):
    world.step(1.0 / 60, 8, 3)
print("READY", flush=True)
sys.stdin.readline()
t0 = time.perf_counter()
for i in range(steps):
    world.step(1.0 / 60, 8, 3)
print("DONE %%.6f %%s" %% (time.perf_counter() - t0, "cext" if b2.Vec2 is not b2.PyVec2 else "pure"), flush=True)
"""


def _reference_processes(ref, cores, settle, steps, warmup, pure, ccd):
    """One pinned process per core, each stepping its own pyramid-20 world through the unmodified reference."""
    procs = [subprocess.Popen([sys.executable, "-c", REF_WORKER % {"ref": ref}, str(c), str(settle), str(steps), str(warmup),
                               "pure" if pure else "cext", "ccd" if ccd else "noccd"],
                              stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True) for c in range(cores)]
    for p in procs:
        assert p.stdout.readline().startswith("READY")
    t0 = time.perf_counter()
    for p in procs:
        p.stdin.write("go\n")
        p.stdin.flush()
    outs = [p.stdout.readline().split() for p in procs]
    wall = time.perf_counter() - t0
    for p in procs:
        p.wait()
    per = sorted(float(o[1]) for o in outs)
    return {"value": N_DYNAMIC * cores * steps / wall, "wall_s": wall, "median_s_per_step": per[len(per) // 2] / steps,
            "backend": outs[0][2], "cores": cores, "steps": steps}


def run_reference_arm(args):
    """`--impl reference`: the UNMODIFIED reference (oracle/_ref, installed with its own C extension by
    __graft_entry__.build) on the host cores, one pinned process per core, each stepping its own settled
    pyramid-20 world through World.step.  Headline = the `build_cext` path with continuous_physics=False (the
    semantics the GPU path implements); the pure-Python path and the script-default CCD-on setting are timed
    after it on a shorter sample and reported beside it.  Falls back to the oracle port when oracle/_ref is absent."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    steps, warmup = args.steps, max(args.warmup, 3)
    extras = None
    if os.path.exists(os.path.join(ref, "pypybox2d", "__init__.py")):
        cores = max(1, min(os.cpu_count() or 1, 64))
        settle = min(args.settle, 120)     # one island from ~step 120; keeps set-up under a minute
        main = _reference_processes(ref, cores, settle, steps, warmup, pure=False, ccd=False)
        value, wall = main["value"], main["wall_s"]
        backend = "C extension (_common, build_cext)" if main["backend"] == "cext" else "pure Python"
        kind, sample = "reference", ("%d processes (one per core, pinned) x %d World.step calls on a settled pyramid-20 "
                                     "world each (%d set-up steps), unmodified pypybox2d from oracle/_ref, %s math "
                                     "types, step(1/60,8,3), continuous_physics=False; median per-core %.3f s/step"
                                     % (cores, steps, settle + warmup, backend, main["median_s_per_step"]))
        ms = 1e3 * wall / steps
        if not args.no_cpu_baseline:
            few = max(2, min(steps, 3))
            ccd_on = _reference_processes(ref, cores, settle, few, 1, pure=False, ccd=True)
            pure_py = _reference_processes(ref, cores, 60, few, 1, pure=True, ccd=False)
            extras = {"cext_ccd_on": {k: ccd_on[k] for k in ("value", "median_s_per_step", "cores", "steps")},
                      "pure_python_ccd_off": dict({k: pure_py[k] for k in ("value", "median_s_per_step", "cores", "steps")},
                                                  note="60 set-up steps (still landing): a lower bound on its cost per step")}
    else:
        cb = cpu_baseline_port(max(5.0, min(args.cpu_seconds, 30.0)), settle=args.settle)
        value, cores, kind, sample = cb["value"], cb["cores"], cb["kind"], cb["sample"]
        ms = 1e3 * N_DYNAMIC * cores / value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "body-steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "pyramid-20 (examples/pyramid-bench.py scene), one world per host core, settled regime",
                       "vel_iters": VEL_ITERS, "pos_iters": POS_ITERS, "dt": "1/60", "continuous_physics": False},
            "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": cores, "kind": kind, "sample": sample,
                             "other_paths": extras},
            "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def pin_to_gpu_numa_node(device_index):
    """Run this process (and first-touch its pinned host buffers) on the NUMA node the GPU hangs off: eight ranks that
    each move ~1 GB per step through pinned memory must not all sit on one socket.  Best effort; returns a note."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "node %d (%d cpus)" % (node, len(allowed))
        return "node %d not in the allowed cpu set" % node
    except Exception as e:
        return "not pinned (%s)" % type(e).__name__


def build_batch(b2, args, S, world_offset, device):
    """`args.worlds` settled worlds of the scene on `device`: `distinct` distinct worlds (own jitter; the mixed scene also
    draws its own shapes per world) are stepped `settle` steps, then their state is tiled over the
    batch (world k of the shard = distinct world k % distinct)."""
    import torch
    from pypybox2d_b200 import scenes
    distinct = min(args.distinct or S["distinct"], args.worlds)
    if S.get("per_world_shapes"):
        worlds = [scenes.SCENES[args.scene](b2, world=world_offset + k) for k in range(distinct)]
        seed = b2.BatchedWorld.from_worlds(worlds, real=args.real, caps=S["caps"], device=device, threads=args.threads)
        del worlds
    else:
        w = scenes.SCENES[args.scene](b2)
        seed = b2.BatchedWorld.replicate(w, distinct, jitter=scenes.JITTER_AMP, config=S["config"], world_offset=world_offset,
                                         real=args.real, caps=S["caps"], device=device, threads=args.threads)
    seed.step(S["dt"], S["vi"], S["pi"], args.settle)
    torch.cuda.synchronize(device)
    if seed.status(raise_on_overflow=False):
        raise RuntimeError("capacity overflow while settling the seed worlds")
    if distinct == args.worlds:
        return seed
    full = seed.tiled(args.worlds, threads=args.threads)
    del seed
    torch.cuda.empty_cache()
    return full


def invariance_check(b2, dist, rank, world_size, device, real):
    """T4 on hardware, outside the timed region: 64 jittered pyramid worlds (global indices 0..63) sharded over the ranks,
    60 steps from the scripted start, per-world checksum rows gathered over NCCL (`dist.gather_rows`); rank 0 compares
    them with its own one-GPU run of the same 64 indices and with the hash committed under tests/golden/."""
    import hashlib
    import torch
    from pypybox2d_b200 import scenes, schema
    n, steps = 64, 60
    caps = schema.Caps(211, 211, 640, npair=1280)
    a, b = dist.shard_range(n, rank, world_size)
    mine = b2.BatchedWorld.replicate(scenes.pyramid(b2), b - a, jitter=scenes.JITTER_AMP, config=3, world_offset=a, real=real,
                                     caps=caps, device=device)
    mine.step(DT, VEL_ITERS, POS_ITERS, steps)
    rows = dist.gather_rows(mine.checksum(), n, rank, world_size)
    if rank != 0:
        return None
    if world_size > 1:
        ref = b2.BatchedWorld.replicate(scenes.pyramid(b2), n, jitter=scenes.JITTER_AMP, config=3, world_offset=0, real=real,
                                        caps=caps, device=device)
        ref.step(DT, VEL_ITERS, POS_ITERS, steps)
        ref_rows = ref.checksum()
    else:
        ref_rows = rows
    same = bool(torch.equal(rows, ref_rows))
    h = hashlib.sha256(rows[:, 0].to(torch.int64).cpu().numpy().tobytes()).hexdigest()[:16]
    golden, golden_ok = None, None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "invariance_pyramid64.json")) as f:
            golden = json.load(f).get(real)
        golden_ok = (golden == h) if golden else None
    except Exception:
        pass
    return {"worlds": n, "steps": steps, "ranks": world_size, "equals_one_gpu_run": same, "hash": h, "golden_hash": golden,
            "golden_ok": golden_ok, "transport": "nccl all_gather" if world_size > 1 else "single process"}


def reference_sample(seconds):
    """cpu_baseline leg of the GPU arm: the UNMODIFIED reference (oracle/_ref, its own C extension) on this box's host
    cores for a bounded sample -- one pinned process per core, each settles its own pyramid-20 world (120 steps: one
    island) and then steps it `steps` times.  None when oracle/_ref is absent."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "pypybox2d", "__init__.py")):
        return None
    cores = max(1, min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 64))
    steps = max(3, min(40, int(seconds / 0.05 / 4)))
    r = _reference_processes(ref, cores, 120, steps, 1, pure=False, ccd=False)
    per_core = N_DYNAMIC / r["median_s_per_step"]
    return {"value": r["value"], "unit": "body-steps/s", "cores": cores, "kind": "reference", "per_core": per_core,
            "sample": "%d processes (one per core, pinned) x %d World.step calls on a settled pyramid-20 world each (121 set-up "
                      "steps), unmodified pypybox2d from oracle/_ref with its C extension, step(1/60,8,3), continuous_physics="
                      "False; median per-core %.3f s/step" % (cores, steps, r["median_s_per_step"])}


def run_b200_arm(args):
    import torch
    import pypybox2d_b200 as b2
    from pypybox2d_b200 import dist, schema

    S = dict(SCENE_TABLE[args.scene])
    if args.worlds <= 0:
        args.worlds = S["worlds"]
    rank, local_rank, world_size = dist.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the step has no CPU fallback (use --impl reference for the CPU arm)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    all_cpus = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa = pin_to_gpu_numa_node(device.index)
    n_total = args.worlds * world_size
    n_dyn, dt, vi, pi = S["n_dyn"], S["dt"], S["vi"], S["pi"]
    bw = build_batch(b2, args, S, rank * args.worlds, device)
    caps = bw.caps
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident timing: K steps, state in HBM (inputs >> L2: no flush needed), one event per step
    bw.step(dt, vi, pi, W)
    torch.cuda.synchronize(device)
    dist.barrier()
    launches0 = bw.info()["launches"]
    sampler = ClockSampler(device.index) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    torch.cuda.synchronize(device)
    ev[0].record()
    for k in range(K):
        bw.step(dt, vi, pi, 1)
        ev[k + 1].record()
    torch.cuda.synchronize(device)
    dist.barrier()
    ms = dist.max_over_ranks(ev[0].elapsed_time(ev[K]), device)
    per_step = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(K))
    clocks = sampler.stop() if sampler else None
    launches = bw.info()["launches"] - launches0
    value = n_dyn * n_total * K / (ms * 1e-3)
    # ---- the same K steps once more with a CUDA event between every two launches: per-kernel durations for the roofline.
    # (A separate pass: with the events in, every kernel runs alone on one stream; the headline pass lets the tails of the
    # solve kernels of neighbouring chunks of worlds overlap on two streams, b2b_api.cu.)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    bw.step(dt, vi, pi, K, time_kernels=True)
    k1.record()
    torch.cuda.synchronize(device)
    serial_ms = k0.elapsed_time(k1) / K
    kms = bw.kernel_ms()

    # ---- end to end through the host-buffer entry point: pinned H2D forces + step + D2H body state, per step
    Ke = args.e2e_steps or K
    rd = torch.float64 if args.real == "f64" else torch.float32
    h_in = torch.zeros((args.worlds, caps.nb, 3), dtype=rd).pin_memory()
    h_out = torch.zeros((args.worlds, caps.nb, 6), dtype=rd).pin_memory()
    bw.step_host(h_in, h_out, dt, vi, pi)           # warm the staging buffers
    dist.barrier()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(Ke):
        bw.step_host(h_in, h_out, dt, vi, pi)
    torch.cuda.synchronize(device)
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0, device)
    e2e_value = n_dyn * n_total * Ke / e2e_s
    status = bw.status(raise_on_overflow=False)
    if status:
        raise RuntimeError("a world overflowed a capacity during the timed steps (status %d): the numbers are invalid" % status)
    launch_info = bw.info()
    h2d_bytes, d2h_bytes = int(h_in.numel() * h_in.element_size()), int(h_out.numel() * h_out.element_size())
    state_gb = sum(t.numel() * t.element_size() for t in bw.t.values()) / 1e9

    # ---- roofline.  Whole step: SURVEY §8d's algorithmic bytes per world-step, 96 Nb + 16 Np + 160 Nc.  Dominant kernel
    # (the one with the largest CUDA-event time in THIS run): the solver stage's share of those bytes -- every body's state
    # read + written and constants read, 96 B per body; every touching contact's constraint read and impulses written,
    # 96 B per contact -- is charged in full to it when it is a solve kernel; a flat kernel is charged its own
    # row of DESIGN.md's table.  Durations are this run's CUDA events between the launches (b2b_kernel_ms).
    D = bw.download(["w_count", "w_island"], worlds=slice(0, min(args.worlds, 4096)))
    nb_mean, np_mean = float(D["w_count"][:, 0].mean()), float(D["w_count"][:, 1].mean())
    nc_mean = float(D["w_count"][:, 2].mean())
    nt_mean = float(D["w_island"][:, 2].mean())
    pos_iters_used = float(D["w_island"][:, 3].mean())
    rb = 8 if args.real == "f64" else 4
    step_bytes = (96 * nb_mean + 16 * np_mean + 160 * nc_mean) * (rb / 4.0)
    solve_bytes = (96 * nb_mean + 96 * nt_mean) * (rb / 4.0)
    kernel_bytes = {"collide": 160 * nc_mean * (rb / 4.0), "plan": (8 * nc_mean + 32 * nb_mean), "prepare": solve_bytes,
                    "solve_velocity": solve_bytes, "solve_position": solve_bytes, "sync": (96 * nb_mean + 32 * np_mean) * (rb / 4.0),
                    "pairs": 16 * np_mean * (rb / 4.0)}
    per_kernel_ms = {k: kms[k] / max(kms["steps"], 1) for k in bw.KERNELS}
    top = max(per_kernel_ms, key=per_kernel_ms.get)
    top_ms = per_kernel_ms[top]
    achieved = kernel_bytes[top] * args.worlds / (top_ms * 1e-3) / 1e9              # GB/s on this rank's GPU
    step_achieved = step_bytes * args.worlds * K / (ms * 1e-3) / 1e9
    # FP32 (non-tensor) view: SURVEY §8d's algorithmic FLOPs per world-step against SMs x 128 lanes x 2 x the clock seen
    npts = 2.0 * nt_mean
    V_, P_ = vi, max(pos_iters_used, 1.0)
    flops_per_world_step = 400 * nc_mean + nt_mean * (270 + 165 * V_) + 150 * npts * P_ + 30 * nb_mean + 100 * np_mean
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(device).multi_processor_count
    fp32_peak = n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12           # TFLOP/s
    fp32_achieved = flops_per_world_step * args.worlds * K / (ms * 1e-3) / 1e12
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    # DRAM traffic of the dominant kernel: only from an ncu capture made by tools/ncu_traffic.py on THIS scene and world count
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        row = tj.get("%s/%d/%s" % (args.scene, args.worlds, top))
        if row:
            traffic, traffic_src = row["dram_bytes_per_launch"], row.get("how")
    except Exception:
        pass

    inv = invariance_check(b2, dist, rank, world_size, device, args.real) if not args.no_invariance else None

    # ---- secondary, reported under config.continuous_on: the same settled batch with continuous collision switched ON
    # (World(continuous_physics=True), the reference's default; World.solve_toi runs in one more kernel after the pairs: in a
    # settled pyramid every step computes the time of impact of the 20 ground contacts and finds no event).  Rank 0 only.
    ccd_on = None
    if rank == 0 and args.ccd_steps > 0:
        bw.set_continuous(True)
        bw.step(dt, vi, pi, 2)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        c0.record()
        bw.step(dt, vi, pi, args.ccd_steps)
        c1.record()
        torch.cuda.synchronize(device)
        ccd_ms = c0.elapsed_time(c1) / args.ccd_steps
        ccd_on = {"value": n_dyn * args.worlds / (ccd_ms * 1e-3), "unit": "body-steps/s", "ms_per_step": ccd_ms, "steps": args.ccd_steps,
                  "worlds": args.worlds, "status": bw.status(raise_on_overflow=False),
                  "what": "the settled batch of the headline with continuous_physics=True on every world, one GPU, state resident"}
        bw.set_continuous(False)

    # ---- secondary, reported under config.from_t0: the WHOLE job of the reference's bench script -- 600 steps from
    # the scripted start, free fall and landing transient included (rank 0 only).  Not the headline: BASELINE's metric
    # is quoted on the settled regime, where a long-running batch spends its time.
    from_t0 = None
    if rank == 0 and args.t0_worlds > 0 and args.t0_steps > 0 and args.scene == "pyramid":
        from pypybox2d_b200 import scenes
        del bw, h_in, h_out
        torch.cuda.empty_cache()
        cold = b2.BatchedWorld.replicate(scenes.pyramid(b2), args.t0_worlds, jitter=scenes.JITTER_AMP, config=3, real=args.real,
                                         caps=S["caps"], device=device, threads=args.threads)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        c0.record()
        cold.step(dt, vi, pi, args.t0_steps)
        c1.record()
        torch.cuda.synchronize(device)
        cold_ms = c0.elapsed_time(c1)
        from_t0 = {"value": n_dyn * args.t0_worlds * args.t0_steps / (cold_ms * 1e-3), "unit": "body-steps/s",
                   "worlds": args.t0_worlds, "steps": args.t0_steps, "ms_per_step": cold_ms / args.t0_steps,
                   "status": cold.status(raise_on_overflow=False),
                   "what": "600 steps from the scripted start (free fall, landing, settling), one GPU, state resident"}
        del cold
        if args.ccd_steps > 0:
            # the reference's bench script verbatim: World((0, -10), True) has continuous collision ON while the pile lands
            # (hundreds of TOI events per world, each a sub-island solve + proxy synchronisation + pair finding)
            torch.cuda.empty_cache()
            cold = b2.BatchedWorld.replicate(scenes.pyramid(b2, continuous_physics=True), args.t0_worlds, jitter=scenes.JITTER_AMP, config=3,
                                             real=args.real, caps=S["caps"], device=device, threads=args.threads)
            torch.cuda.synchronize(device)
            c0.record()
            cold.step(dt, vi, pi, args.t0_steps)
            c1.record()
            torch.cuda.synchronize(device)
            cold_ms = c0.elapsed_time(c1)
            from_t0["continuous_on"] = {"value": n_dyn * args.t0_worlds * args.t0_steps / (cold_ms * 1e-3), "unit": "body-steps/s",
                                        "ms_per_step": cold_ms / args.t0_steps, "status": cold.status(raise_on_overflow=False)}
            del cold

    if world_size > 1:
        import torch.distributed as tdist
        dist.barrier()
        if tdist.is_initialized():
            tdist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC if args.scene == "pyramid" else "body-steps/sec (batched %s worlds)" % args.scene,
        "value": value, "unit": "body-steps/s", "n_gpus": world_size, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "ms_per_step_min_median_max": [per_step[0], per_step[len(per_step) // 2], per_step[-1]],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.real, "data": "synthetic",
        "config": {"workload": "%s x %d worlds per GPU (%d total), settled regime" % (S["what"], args.worlds, n_total),
                   "scene": args.scene, "dynamic_bodies_per_world": n_dyn,
                   "vel_iters": vi, "pos_iters": pi, "dt": S["dt_text"], "continuous_physics": False,
                   "distinct_worlds_per_gpu": min(args.distinct or S["distinct"], args.worlds), "settle_steps": args.settle,
                   "contacts_per_world": nc_mean, "touching_contacts_per_world": nt_mean,
                   "position_iterations_used": pos_iters_used, "l2": "inputs larger than L2 (%.1f GB of state per GPU), no flush"
                                                        % (state_gb,),
                   "parallelism": "worlds sharded by contiguous index blocks, one process per GPU, no collective on the step path",
                   "host_numa": numa, "launch": launch_info, "status": status, "from_t0": from_t0, "continuous_on": ccd_on},
        "e2e": {"value": e2e_value, "unit": "body-steps/s", "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke,
                "h2d_bytes_per_step": h2d_bytes * world_size, "d2h_bytes_per_step": d2h_bytes * world_size,
                "ratio_to_value": e2e_value / value,
                "how": "b2b_step_host: forces in / narrowphase / solve / state out pipelined over chunks of worlds on two copy streams"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "invariance_ok": (None if inv is None else bool(inv["equals_one_gpu_run"] and inv["golden_ok"] is not False)),
        "invariance": inv,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "traffic_gbs": (None if not traffic else traffic / (top_ms * 1e-3) / 1e9),
                     "traffic_frac": (None if not traffic else traffic / (top_ms * 1e-3) / 1e9 / peak),
                     "kernel": "b2b::%s_kernel_%s" % (top + (("_revolute" if bw._caps_struct.joint_types == 2 else "_joints") if caps.nj > 0 and top.startswith("solve_") else ""), args.real),
                     "kernel_ms_per_launch": top_ms, "kernel_share_of_step": top_ms / max(sum(per_kernel_ms.values()), 1e-9),
                     "algorithmic_bytes_per_world_step": kernel_bytes[top],
                     "kernel_ms": per_kernel_ms, "ms_per_step_one_stream_with_events": serial_ms,
                     "whole_step_frac": step_achieved / peak,
                     "step": {"achieved": step_achieved, "frac": step_achieved / peak, "algorithmic_bytes_per_world_step": step_bytes,
                              "formula": "96 Nb + 16 Np + 160 Nc (SURVEY 8d)"},
                     "fp32": {"achieved_tflops": fp32_achieved, "peak_tflops": fp32_peak, "frac": fp32_achieved / fp32_peak,
                              "algorithmic_flops_per_world_step": flops_per_world_step,
                              "peak_source": "%d SMs x 128 lanes x 2 x %.0f MHz (clock sampled during the run)" % (n_sm, sm_mhz)},
                     "note": "`frac` counts ALGORITHMIC bytes; `traffic_frac` is the DRAM traffic the kernel really moves (ncu, same "
                             "command) over this run's kernel time: the constraint records are re-streamed from HBM by every sweep, "
                             "so the kernel sits at about two thirds of the HBM peak while doing ~10x the algorithmic bytes "
                             "(DESIGN.md 4, 6)"},
    }
    if not args.no_cpu_baseline and world_size == 1 and args.scene == "pyramid":
        if all_cpus:
            os.sched_setaffinity(0, all_cpus)            # the CPU legs use every host core again
        port = cpu_baseline_port(min(args.cpu_seconds, 10.0), settle=args.settle)
        refb = reference_sample(args.cpu_seconds)
        if refb is not None:
            line["cpu_baseline"] = dict(refb, port=port)
            line["speedup_vs_cext_per_core"] = value / refb["per_core"]
            line["speedup_vs_cext_all_cores"] = value / refb["value"]
            line["e2e_speedup_vs_cext_all_cores"] = e2e_value / refb["value"]
        else:
            line["cpu_baseline"] = port
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))


def main():
    # rank 0's stdout carries exactly one JSON line: keep NCCL's own chatter (version banner, INFO logs) off it
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    args = parse()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank == 0:
            run_reference_arm(args)
        return
    global SCENE_TABLE
    SCENE_TABLE = _scene_table()
    run_b200_arm(args)


if __name__ == "__main__":
    main()
