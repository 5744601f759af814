"""`BatchedWorld`: a batch of independent worlds resident on one GPU.

Host-side mirror of the reference's `World` for the step path only: it owns the device
structure-of-arrays buffers (torch tensors, `schema.FIELDS`), hands their pointers to libb2b.so
(`include/b2b.h`) and launches `World.step` for every world of the batch at once.  No per-body
Python runs on the step path; no CPU implementation of the step exists here -- without the CUDA
library or a GPU every stepping call raises.
"""
import ctypes

import numpy as np

from . import schema
from . import _lib

_TORCH_DTYPES = None


def _torch():
    import torch
    global _TORCH_DTYPES
    if _TORCH_DTYPES is None:
        _TORCH_DTYPES = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                         np.dtype(np.int32): torch.int32, np.dtype(np.uint32): torch.int32,
                         np.dtype(np.uint8): torch.uint8}
    return torch


def world_caps(worlds, contacts_per_body=4, slack=64):
    """Capacities for a batch of host worlds.  Contacts are between fixtures, so a compound body with many fixtures
    needs rows per FIXTURE, not per body (`World.step` raises if a step still runs out)."""
    nb = max(len(w.bodies) for w in worlds)
    nf = max(sum(f._shape.child_count for b in w.bodies for f in b.fixtures) for w in worlds)    # proxies: one per shape child
    nj = max(len(getattr(w, "joints", ())) for w in worlds)
    return schema.Caps(nb, nf, contacts_per_body * max(nb, nf) + slack, nj=nj)


class BatchedWorld(object):
    def __init__(self, arrays, caps, real="f32", device=None, threads=0, ctas_per_sm=0, lanes_aux=0):
        """`arrays`: host numpy dict laid out as `schema.alloc` (any real dtype; cast to `real`)."""
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.B2BError("no CUDA device: the batched step has no CPU fallback")
        self.real = real
        self.caps = caps
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n_worlds = int(arrays["w_flags"].shape[0])
        self.n_shapes = int(arrays["s_meta"].shape[0])
        self.inv_dt0 = 0.0
        self.t = {}
        for name, scope, tail, kind in schema.FIELDS:
            want = schema.np_dtype(kind, real)
            host = np.ascontiguousarray(arrays[name]).astype(want, copy=False)
            self.t[name] = self._to_device(host)
        self._create_context(threads, ctas_per_sm, lanes_aux)

    def _create_context(self, threads=0, ctas_per_sm=0, lanes_aux=0):
        """b2b_create + b2b_bind for the tensors in self.t."""
        torch = _torch()
        self._ctx = ctypes.c_void_p()
        c = _lib.Caps()
        caps = self.caps
        c.n_worlds, c.nb, c.nf, c.nc, c.npair, c.nl = self.n_worlds, caps.nb, caps.nf, caps.nc, caps.npair, caps.nl
        c.n_shapes = self.n_shapes
        c.real_bytes = 8 if self.real == "f64" else 4
        c.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        c.threads, c.ctas_per_sm, c.reserved = int(threads), int(ctas_per_sm), int(lanes_aux)
        c.nj = int(getattr(caps, "nj", 0))
        c.joint_types = int(self._joint_types()) if c.nj > 0 else 0
        self._caps_struct = c
        _lib.check(_lib.lib().b2b_create(ctypes.byref(c), ctypes.byref(self._ctx)), "b2b_create")
        self._bind()

    def _joint_types(self):
        """Bit t set: some joint of the batch has type t (b2b_caps.joint_types: a batch of revolute joints only runs kernels
        without the type dispatch)."""
        d, n = self.t["j_def"][..., 0], self.t["w_joint"][:, 0:1]
        torch = _torch()
        used = torch.arange(d.shape[1], device=d.device).unsqueeze(0) < n
        types = torch.unique(d[used]).tolist()
        mask = 0
        for t in types:
            mask |= 1 << int(t)
        return mask

    @classmethod
    def from_tensors(cls, tensors, caps, real="f32", inv_dt0=0.0, **kw):
        """Wrap existing device tensors (one per schema field, already in the working dtypes) in a new context."""
        torch = _torch()
        self = cls.__new__(cls)
        self.real, self.caps = real, caps
        self.t = dict(tensors)
        self.device = self.t["w_flags"].device
        self.n_worlds, self.n_shapes = int(self.t["w_flags"].shape[0]), int(self.t["s_meta"].shape[0])
        self.inv_dt0 = float(inv_dt0)
        self._create_context(kw.get("threads", 0), kw.get("ctas_per_sm", 0), kw.get("lanes_aux", 0))
        return self

    def tiled(self, n_worlds, **kw):
        """A new batch of `n_worlds` worlds whose world k is a copy of this batch's world k % len(self): used to
        spread a set of settled worlds over a large benchmark batch."""
        reps = (n_worlds + self.n_worlds - 1) // self.n_worlds
        t = {}
        for name, scope, tail, kind in schema.FIELDS:
            x = self.t[name]
            t[name] = x if scope == "S" else x.repeat((reps,) + (1,) * (x.dim() - 1))[:n_worlds].contiguous()
        return BatchedWorld.from_tensors(t, self.caps, real=self.real, inv_dt0=self.inv_dt0, **kw)

    # -- construction -----------------------------------------------------------------------
    @classmethod
    def from_worlds(cls, worlds, real="f32", caps=None, device=None, **kw):
        caps = caps or world_caps(worlds)
        table = schema.ShapeTable()
        for w in worlds:
            schema.add_shapes(table, w)
        A = schema.alloc(caps, len(worlds), max(len(table), 1), "f64")
        for i, w in enumerate(worlds):
            schema.pack_world(w, A, i, table)
        A["s_meta"], A["s_radius"], A["s_data"] = table.arrays("f64")
        bw = cls(A, caps, real=real, device=device, **kw)
        if real != "f64":
            bw.init_proxies()        # proxies are built in the working precision, exactly as `replicate` builds them
        return bw

    @classmethod
    def replicate(cls, world, n, jitter=0.0, config=0, world_offset=0, real="f32", caps=None, device=None, **kw):
        """`n` copies of `world`; copy k is world index `world_offset + k` of the global batch.  World index 0
        is the scene as scripted; every other index moves each dynamic body by jitter*U(-1,1) in x and y, with
        U from `scenes.u01_torch(config, world index, 2*body + axis)` -- so a world's scene depends only on its
        global index.  The expansion and the jitter run on the device; proxies are (re)built there too
        (`b2b_init_proxies`)."""
        torch = _torch()
        from . import scenes
        caps = caps or world_caps([world])
        table = schema.ShapeTable()
        A1 = schema.alloc(caps, 1, 1, "f64")
        schema.add_shapes(table, world)
        A1 = schema.alloc(caps, 1, max(len(table), 1), "f64")
        schema.pack_world(world, A1, 0, table)
        A1["s_meta"], A1["s_radius"], A1["s_data"] = table.arrays("f64")
        self = cls.__new__(cls)
        if not torch.cuda.is_available():
            raise _lib.B2BError("no CUDA device: the batched step has no CPU fallback")
        self.real, self.caps = real, caps
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n_worlds, self.n_shapes, self.inv_dt0 = int(n), int(A1["s_meta"].shape[0]), 0.0
        self.t = {}
        for name, scope, tail, kind in schema.FIELDS:
            host = np.ascontiguousarray(A1[name])
            if scope == "S":
                self.t[name] = self._to_device(host.astype(schema.np_dtype(kind, real), copy=False))
                continue
            one = self._to_device(host)                      # fp64 / ints, one world
            full = one.expand((n,) + tuple(one.shape[1:]))
            if name in ("b_xf", "b_pos") and jitter != 0.0:
                full = full.clone()                           # jittered in fp64 below, cast after
                self.t[name] = full
            else:
                want = _TORCH_DTYPES[np.dtype(schema.np_dtype(kind, real))]
                self.t[name] = full.to(want).contiguous() if full.dtype != want else full.contiguous()
        if jitter != 0.0:
            nb = caps.nb
            widx = torch.arange(world_offset, world_offset + n, device=self.device, dtype=torch.int64).view(n, 1, 1)
            k = torch.arange(2 * nb, device=self.device, dtype=torch.int64).view(1, nb, 2)
            u = scenes.u01_torch(config, widx, k)
            d = jitter * (2.0 * u - 1.0)
            dyn = ((self.t["b_flags"] & schema.BODY_TYPE_MASK) == 1).unsqueeze(-1)
            d = torch.where(dyn & (widx != 0), d, torch.zeros_like(d))
            xf, pos, mass = self.t["b_xf"], self.t["b_pos"], self.t["b_mass"].to(torch.float64)
            xf[..., 0:2] += d
            # sweep.c = xf * local_center in the reference's operand order (body.py:95)
            lcx, lcy = mass[..., 2], mass[..., 3]
            pos[..., 0] = xf[..., 0] + xf[..., 2] * lcx + (-xf[..., 3]) * lcy
            pos[..., 1] = xf[..., 1] + xf[..., 3] * lcx + xf[..., 2] * lcy
            want = torch.float64 if real == "f64" else torch.float32
            self.t["b_xf"] = xf.to(want).contiguous()
            self.t["b_pos"] = pos.to(want).contiguous()
            self.t["b_pos0"] = self.t["b_pos"].clone()
            self.t["b_pos0"][..., 3] = 0
        self._create_context(kw.get("threads", 0), kw.get("ctas_per_sm", 0), kw.get("lanes_aux", 0))
        if jitter != 0.0 or real != "f64":
            self.init_proxies()      # proxies are built in the working precision, exactly as `from_worlds` builds them
        return self

    def _to_device(self, host):
        torch = _torch()
        if host.dtype == np.uint32:
            host = host.view(np.int32)
        return torch.from_numpy(np.ascontiguousarray(host)).to(self.device)

    def _bind(self):
        b = _lib.Buffers()
        for name in schema.FIELD_NAMES:
            t = self.t[name]
            assert t.is_contiguous(), name
            setattr(b, name, t.data_ptr())
        self._buffers = b
        # the context's island schedule is empty: whatever the arrays say, the first step must rebuild it
        self.t["w_flags"] &= ~schema.WORLD_GRAPH_VALID
        _lib.check(_lib.lib().b2b_bind(self._ctx, ctypes.byref(b)), "b2b_bind")

    def __del__(self):
        try:
            if getattr(self, "_ctx", None):
                _lib.lib().b2b_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    # -- stepping ---------------------------------------------------------------------------
    def _stream(self):
        torch = _torch()
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def init_proxies(self):
        _lib.check(_lib.lib().b2b_init_proxies(self._ctx, self._stream()), "b2b_init_proxies")

    PROFILE_SLOTS = ["pairs0", "collide", "adjacency", "islands", "solve", "sync", "pairs", "clear",
                     "s_stage", "s_integrate", "s_schedule", "s_init", "s_warm", "s_velocity", "s_store", "s_position",
                     "s_writeback", "s_sleep"]

    def step(self, dt, vel_iters, pos_iters, n_steps=1, profile=False, time_kernels=False, group_kernel=False):
        """World.step for every world, `n_steps` times (asynchronous on the current stream)."""
        flags = (1 if profile else 0) | (2 if time_kernels else 0) | (32 if group_kernel else 0)   # B2B_STEP_* (include/b2b.h)
        _lib.check(_lib.lib().b2b_step(self._ctx, float(dt), float(self.inv_dt0), int(vel_iters), int(pos_iters),
                                       int(n_steps), flags, self._stream()), "b2b_step")
        if dt > 0.0 and n_steps > 0:
            self.inv_dt0 = 1.0 / dt

    def solve_islands(self, dt, vel_iters, pos_iters):
        """Test hook: island construction + Island.solve only (no narrowphase before, no proxy synchronisation / new contacts
        after); `step(0, ...)` before it runs the narrowphase alone."""
        _lib.check(_lib.lib().b2b_solve_islands(self._ctx, float(dt), float(self.inv_dt0), int(vel_iters), int(pos_iters), 0,
                                                self._stream()), "b2b_solve_islands")

    @property
    def continuous(self):
        """True when some world of the batch has continuous collision on (World(continuous_physics=True), world.py:78):
        World.solve_toi then runs in the last kernel of every step."""
        if getattr(self, "_continuous", None) is None:
            self._continuous = bool((self.t["w_flags"] & schema.WORLD_CONTINUOUS).any().item())
        return self._continuous

    def set_continuous(self, on):
        """World.continuous_physics = on, for every world of the batch."""
        if on:
            self.t["w_flags"] |= schema.WORLD_CONTINUOUS
        else:
            self.t["w_flags"] &= ~schema.WORLD_CONTINUOUS
        self._continuous = bool(on)

    def step_host(self, h_force_torque, h_state_out, dt, vel_iters, pos_iters, n_steps=1):
        """The end-to-end call: pinned host forces in, pinned host body state out, synchronous."""
        fin = ctypes.c_void_p(h_force_torque.data_ptr()) if h_force_torque is not None else None
        fout = ctypes.c_void_p(h_state_out.data_ptr()) if h_state_out is not None else None
        flags = 128 if self.continuous else 0          # B2B_STEP_CONTINUOUS: the state is read back after World.solve_toi
        _lib.check(_lib.lib().b2b_step_host(self._ctx, fin, fout, float(dt), float(self.inv_dt0), int(vel_iters),
                                            int(pos_iters), int(n_steps), flags, self._stream()), "b2b_step_host")
        if dt > 0.0 and n_steps > 0:
            self.inv_dt0 = 1.0 / dt

    def status(self, raise_on_overflow=True):
        out = ctypes.c_uint32(0)
        rc = _lib.lib().b2b_status(self._ctx, ctypes.byref(out), self._stream())
        _lib.check(rc, "b2b_status", allow_overflow=not raise_on_overflow)
        return int(out.value)

    def checksum(self):
        """[n_worlds, 4] float64 tensor: state hash (the low 52 bits of an FNV-1a over the packed body state, as an exactly
        representable integer: comparable with ==), energy, n_contacts, n_awake (device)."""
        torch = _torch()
        out = torch.empty((self.n_worlds, 4), dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().b2b_checksum(self._ctx, ctypes.c_void_p(out.data_ptr()), self._stream()), "b2b_checksum")
        return out

    def profile(self, reset=True):
        """Per-phase cycles of thread 0 summed over CTAs, for steps run with profile=True (development aid)."""
        out = (ctypes.c_int64 * 32)()
        _lib.check(_lib.lib().b2b_profile(self._ctx, out, 1 if reset else 0), "b2b_profile")
        d = {n: int(out[i]) for i, n in enumerate(self.PROFILE_SLOTS)}
        d["world_steps"] = int(out[31])
        return d

    KERNELS = ["collide", "plan", "prepare", "solve_velocity", "solve_position", "sync", "pairs"]

    def kernel_ms(self):
        """Total ms per kernel (KERNELS) of the steps run with time_kernels=True since the last call, from CUDA events
        between the launches; `solve` = prepare + solve_velocity + solve_position, `broadphase` = sync + pairs."""
        out = (ctypes.c_float * 8)()
        n = ctypes.c_int32(0)
        _lib.check(_lib.lib().b2b_kernel_ms(self._ctx, out, ctypes.byref(n)), "b2b_kernel_ms")
        d = {k: float(out[i]) for i, k in enumerate(self.KERNELS)}
        d["solve"] = d["prepare"] + d["solve_velocity"] + d["solve_position"]
        d["broadphase"] = d["sync"] + d["pairs"]
        d["total"] = float(out[7])
        d["steps"] = int(n.value)
        return d

    def info(self):
        g, t, s, n = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
        _lib.check(_lib.lib().b2b_info(self._ctx, ctypes.byref(g), ctypes.byref(t), ctypes.byref(s), ctypes.byref(n)),
                   "b2b_info")
        return {"grid": g.value, "threads": t.value, "smem_bytes": s.value, "launches": n.value}

    def collide_pairs(self, shape_a, shape_b, xf_a, xf_b):
        """Test hook: narrowphase on explicit pairs -> (mtype, mf, pts, id) numpy arrays."""
        torch = _torch()
        rd = torch.float64 if self.real == "f64" else torch.float32
        n = len(shape_a)
        sa = torch.as_tensor(np.asarray(shape_a, np.int32), device=self.device)
        sb = torch.as_tensor(np.asarray(shape_b, np.int32), device=self.device)
        xa = torch.as_tensor(np.asarray(xf_a, np.float64), device=self.device).to(rd).contiguous()
        xb = torch.as_tensor(np.asarray(xf_b, np.float64), device=self.device).to(rd).contiguous()
        mtype = torch.zeros(n, dtype=torch.int32, device=self.device)
        mf = torch.zeros((n, 4), dtype=rd, device=self.device)
        pts = torch.zeros((n, 4), dtype=rd, device=self.device)
        ids = torch.zeros((n, 2), dtype=torch.int32, device=self.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().b2b_collide_pairs(self._ctx, n, p(sa), p(sb), p(xa), p(xb), p(mtype), p(mf), p(pts), p(ids),
                                                self._stream()), "b2b_collide_pairs")
        torch.cuda.synchronize(self.device)
        return mtype.cpu().numpy(), mf.cpu().numpy(), pts.cpu().numpy(), ids.cpu().numpy().view(np.uint32)

    # -- queries (World.query_aabb / ray_cast, Fixture.test_point; include/b2b.h) -----------------------------
    def _query_inputs(self, worlds, reals, width):
        torch = _torch()
        rd = torch.float64 if self.real == "f64" else torch.float32
        w = torch.as_tensor(np.ascontiguousarray(worlds, dtype=np.int32), device=self.device)
        x = torch.as_tensor(np.ascontiguousarray(reals, dtype=np.float64).reshape(-1, width), device=self.device).to(rd).contiguous()
        if w.numel() != x.shape[0]:
            raise ValueError("one world index per query")
        return w, x, rd

    def query_aabb(self, worlds, boxes, max_hits=64):
        """World.query_aabb for a batch of queries: worlds[q] (int), boxes[q] = (lx, ly, ux, uy).  Returns (counts[q], fixtures[q,
        max_hits]) as numpy arrays: the number of fixtures whose fat AABB overlaps the box and the first `max_hits` of them in
        ascending fixture order (-1 padding)."""
        torch = _torch()
        w, box, rd = self._query_inputs(worlds, boxes, 4)
        n = int(w.numel())
        count = torch.zeros(n, dtype=torch.int32, device=self.device)
        fix = torch.full((n, max_hits), -1, dtype=torch.int32, device=self.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().b2b_query_aabb(self._ctx, n, p(w), p(box), int(max_hits), p(count), p(fix), self._stream()), "b2b_query_aabb")
        torch.cuda.synchronize(self.device)
        return count.cpu().numpy(), fix.cpu().numpy()

    def ray_cast(self, worlds, segments, closest=False, max_hits=64):
        """World.ray_cast for a batch of rays: segments[q] = (x1, y1, x2, y2).  Returns (counts[q], fixtures[q, max_hits],
        hits[q, max_hits, 5] = fraction, normal.x, normal.y, point.x, point.y): every fixture hit in ascending fixture order, or
        with closest=True only the closest one in slot 0.  counts[q] = -1 for a zero-length segment."""
        torch = _torch()
        w, seg, rd = self._query_inputs(worlds, segments, 4)
        n = int(w.numel())
        count = torch.zeros(n, dtype=torch.int32, device=self.device)
        fix = torch.full((n, max_hits), -1, dtype=torch.int32, device=self.device)
        hit = torch.zeros((n, max_hits, 5), dtype=rd, device=self.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().b2b_ray_cast(self._ctx, n, p(w), p(seg), 1 if closest else 0, int(max_hits), p(count), p(fix), p(hit),
                                           self._stream()), "b2b_ray_cast")
        torch.cuda.synchronize(self.device)
        return count.cpu().numpy(), fix.cpu().numpy(), hit.cpu().numpy().astype(np.float64)

    def test_points(self, worlds, fixtures, points):
        """Fixture.test_point for a batch of (world, fixture, point) triples -> bool array."""
        torch = _torch()
        w, pt, rd = self._query_inputs(worlds, points, 2)
        f = torch.as_tensor(np.ascontiguousarray(fixtures, dtype=np.int32), device=self.device)
        n = int(w.numel())
        out = torch.zeros(n, dtype=torch.uint8, device=self.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().b2b_test_points(self._ctx, n, p(w), p(f), p(pt), p(out), self._stream()), "b2b_test_points")
        torch.cuda.synchronize(self.device)
        return out.cpu().numpy().astype(bool)

    # -- transfers --------------------------------------------------------------------------
    def download(self, names=None, worlds=None):
        """Host numpy copies (uint32 fields restored) of the named fields, optionally for a slice of worlds."""
        _torch().cuda.synchronize(self.device)
        out = {}
        for name, scope, tail, kind in schema.FIELDS:
            if names is not None and name not in names:
                continue
            t = self.t[name]
            if worlds is not None and scope != "S":
                t = t[worlds]
            a = t.cpu().numpy()
            if kind == "u32":
                a = a.view(np.uint32)
            out[name] = a
        return out

    def upload(self, arrays, worlds=None):
        """Overwrite device fields from host arrays (teacher forcing / settled starts)."""
        torch = _torch()
        for name, a in arrays.items():
            kind = {f[0]: f[3] for f in schema.FIELDS}[name]
            scope = {f[0]: f[1] for f in schema.FIELDS}[name]
            host = np.ascontiguousarray(a).astype(schema.np_dtype(kind, self.real), copy=False)
            if host.dtype == np.uint32:
                host = host.view(np.int32)
            src = torch.from_numpy(host).to(self.device)
            if worlds is not None and scope != "S":
                self.t[name][worlds].copy_(src)
            else:
                self.t[name].copy_(src)
        if "j_def" in arrays and self.caps.nj > 0 and (self._joint_types() & ~int(self._caps_struct.joint_types)):
            raise _lib.B2BError("upload: a joint type that this batch was not created with (b2b_caps.joint_types); create a new batch")
        self.t["w_flags"] &= ~schema.WORLD_GRAPH_VALID      # host-side mutation: rebuild the island schedule
        self.t["w_flags"] |= schema.WORLD_TOI_DIRTY         # ... and let the next World.solve_toi reset whatever alpha0 came in
        self._continuous = None                             # (w_flags may have been replaced)

    # -- single-world facade support ----------------------------------------------------------
    def upload_bodies_from_world(self, world, w=0):
        A = schema.alloc(self.caps, 1, 1, "f64")
        table = schema.ShapeTable()
        schema.pack_world(world, A, 0, table)
        names = ["b_xf", "b_pos", "b_pos0", "b_vel", "b_force", "b_flags", "b_mass", "b_damp", "f_mat", "w_gravity"]
        if getattr(world, "_joints_dirty", False):
            # a joint's parameters or flags were set between steps (joints.py setters); the joint list itself only changes
            # through World.add_joint / destroy_joint, which re-pack the world
            names += ["j_def", "j_par", "j_imp"]
            world._joints_dirty = False
        self.upload({k: A[k] for k in names}, worlds=slice(w, w + 1))

    def download_into_world(self, world, w=0):
        from .common import Vec2
        D = self.download(["b_xf", "b_pos", "b_vel", "b_pos0", "b_force", "b_flags", "b_island"], worlds=slice(w, w + 1))
        for i, b in enumerate(world.bodies):
            xf = D["b_xf"][0, i]
            b._xf = (float(xf[0]), float(xf[1]), float(xf[2]), float(xf[3]))
            p, v, p0 = D["b_pos"][0, i], D["b_vel"][0, i], D["b_pos0"][0, i]
            b._c = Vec2(p[0], p[1]); b._a = float(p[2]); b._sleep_time = float(p[3])
            b._c0 = Vec2(p0[0], p0[1]); b._a0 = float(p0[2])
            b._linear_velocity = Vec2(v[0], v[1]); b._angular_velocity = float(v[2]); b._torque = float(v[3])
            b._force = Vec2(*D["b_force"][0, i])
            b._awake = bool(int(D["b_flags"][0, i]) & schema.BODY_AWAKE)
            isl = int(D["b_island"][0, i])
            b._island_index = isl if isl >= 0 else b._island_index
        if getattr(world, "joints", None):
            J = self.download(["j_def", "j_imp", "j_tmp"], worlds=slice(w, w + 1))
            for k, j in enumerate(world.joints):
                j._read_back(J["j_def"][0, k, 3], J["j_imp"][0, k])
                if j._type in (schema.JOINT_DISTANCE, schema.JOINT_ROPE):
                    j._u = (float(J["j_tmp"][0, k, 4]), float(J["j_tmp"][0, k, 5]))
                elif j._type == schema.JOINT_PRISMATIC:
                    t = J["j_tmp"][0, k]
                    j._axis_perp = ((float(t[4]), float(t[5])), (float(t[6]), float(t[7])))

    def contact_views(self, world, w=0):
        from .world import ContactView
        from .common import Vec2
        D = self.download(["w_count", "c_fix", "c_flags", "c_mat", "c_mtype", "c_mf", "c_pts", "c_imp"],
                          worlds=slice(w, w + 1))
        fixtures = [f for f, c in schema.world_proxies(world)]      # device fixture rows are proxies: (fixture, child)
        out = []
        for k in range(int(D["w_count"][0, 2])):
            c = ContactView()
            c.fixture_a, c.fixture_b = fixtures[int(D["c_fix"][0, k, 0])], fixtures[int(D["c_fix"][0, k, 1])]
            fl = int(D["c_flags"][0, k])
            c.touching, c.enabled = bool(fl & schema.CONTACT_TOUCHING), bool(fl & schema.CONTACT_ENABLED)
            c.friction, c.restitution = float(D["c_mat"][0, k, 0]), float(D["c_mat"][0, k, 1])
            mt = int(D["c_mtype"][0, k])
            c.manifold_type, c.point_count = mt & 0xff, (mt >> 8) & 0xff
            mf, pts, imp = D["c_mf"][0, k], D["c_pts"][0, k], D["c_imp"][0, k]
            c.local_normal, c.local_point = Vec2(mf[0], mf[1]), Vec2(mf[2], mf[3])
            c.points = [(Vec2(pts[2 * j], pts[2 * j + 1]), float(imp[2 * j]), float(imp[2 * j + 1]))
                        for j in range(c.point_count)]
            out.append(c)
        return out
