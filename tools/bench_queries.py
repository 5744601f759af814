"""Throughput of the query kernels on the bench batch (development aid): N rays / boxes spread over a settled pyramid batch,
device arrays in and out, CUDA events.  Algorithmic bytes per query: the world's proxy boxes (16 B x Np) + for a ray the
transform and shape rows of the proxies whose box the ray crosses."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pypybox2d_b200 as b2
from pypybox2d_b200 import scenes, schema, _lib
worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
caps = schema.Caps(211, 211, 640, npair=1280)
seed = b2.BatchedWorld.replicate(scenes.pyramid(b2), 1024, jitter=1e-2, config=3, real="f32", caps=caps)
seed.step(1 / 60, 8, 3, 200)
bw = seed.tiled(worlds)
dev = bw.device
g = torch.Generator(device=dev); g.manual_seed(1)
w = torch.randint(0, worlds, (nq,), device=dev, dtype=torch.int32, generator=g)
c = torch.rand((nq, 2), device=dev, generator=g) * torch.tensor([24.0, 24.0], device=dev) + torch.tensor([-8.0, 0.0], device=dev)
h = torch.rand((nq, 2), device=dev, generator=g) * 1.5 + 0.1
box = torch.cat([c - h, c + h], 1).contiguous()
seg = torch.cat([c, c + (torch.rand((nq, 2), device=dev, generator=g) - 0.5) * 20.0], 1).contiguous()
H = 16
cnt = torch.zeros(nq, dtype=torch.int32, device=dev); fix = torch.zeros((nq, H), dtype=torch.int32, device=dev)
hit = torch.zeros((nq, H, 5), dtype=torch.float32, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
L, st = _lib.lib(), bw._stream()
out = {}
for name, call in (("query_aabb", lambda: L.b2b_query_aabb(bw._ctx, nq, p(w), p(box), H, p(cnt), p(fix), st)),
                   ("ray_cast_all", lambda: L.b2b_ray_cast(bw._ctx, nq, p(w), p(seg), 0, H, p(cnt), p(fix), p(hit), st)),
                   ("ray_cast_closest", lambda: L.b2b_ray_cast(bw._ctx, nq, p(w), p(seg), 1, H, p(cnt), p(fix), p(hit), st))):
    for _ in range(2):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = nq * 211 * 16 / 1e9
    out[name] = {"queries": nq, "worlds": worlds, "ms": ms, "queries_per_s": nq / (ms * 1e-3), "proxy_box_GBps": gb / (ms * 1e-3),
                 "mean_hits": float(cnt.float().clamp(min=0).mean())}
print(json.dumps(out))
