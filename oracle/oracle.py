"""ctypes binding of the C oracle (TEST INFRASTRUCTURE ONLY -- see oracle/b2_oracle.h).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  Operates in place on fp64 host arrays laid out as pypybox2d_b200.schema.FIELDS.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FIELD_ORDER = ("w_count,w_flags,w_status,w_gravity,w_island,"
               "b_xf,b_pos,b_vel,b_pos0,b_force,b_mass,b_damp,b_flags,b_island,"
               "f_body,f_shape,f_mat,f_filter,f_fat,f_tight,f_moved,"
               "c_fix,c_flags,c_mat,c_mtype,c_mf,c_pts,c_imp,c_id,"
               "p_pairs,l_body,l_contact,l_start,s_meta,s_radius,s_data,"
               "w_joint,j_def,j_par,j_imp,j_tmp,j_adj,b_jadj,l_joint,l_jstart").split(",")


class OrcWorld(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int32) for n in ("nb_cap", "nf_cap", "nc_cap", "npair_cap", "nl_cap", "n_shapes", "nj_cap")] +
                [(n, ctypes.c_void_p) for n in FIELD_ORDER])


def build(force=False):
    """Compile oracle/libb2oracle.so with the committed Makefile (gcc, no network, seconds)."""
    so = os.path.join(_HERE, "libb2oracle.so")
    src = os.path.join(_HERE, "b2_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libb2oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_step.argtypes = [ctypes.POINTER(OrcWorld), ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.orc_step.restype = ctypes.c_int
        L.orc_run.argtypes = [ctypes.POINTER(OrcWorld), ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                              ctypes.c_int]
        L.orc_run.restype = ctypes.c_int
        L.orc_collide_pair.argtypes = [ctypes.POINTER(OrcWorld), ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]
        L.orc_collide_pair.restype = ctypes.c_int
        L.orc_query_aabb.argtypes = [ctypes.POINTER(OrcWorld), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_query_aabb.restype = ctypes.c_int
        L.orc_ray_cast.argtypes = [ctypes.POINTER(OrcWorld), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_ray_cast.restype = ctypes.c_int
        L.orc_test_point.argtypes = [ctypes.POINTER(OrcWorld), ctypes.c_int, ctypes.c_void_p]
        L.orc_test_point.restype = ctypes.c_int
        L.orc_field_names.restype = ctypes.c_char_p
        _LIB = L
    return _LIB


class Oracle(object):
    """One world (row `w`) of a batch of fp64 schema arrays, stepped in place by the C oracle."""

    def __init__(self, arrays, w=0):
        self.arrays = arrays
        self.w = w
        self.inv_dt0 = 0.0
        A = arrays
        for name in FIELD_ORDER:
            a = A[name]
            if not a.flags["C_CONTIGUOUS"]:
                raise ValueError("%s must be C-contiguous" % name)
            if a.dtype == np.float32:
                raise ValueError("the oracle runs on fp64 arrays")
        s = OrcWorld()
        s.nb_cap = A["b_flags"].shape[1]
        s.nf_cap = A["f_body"].shape[1]
        s.nc_cap = A["c_flags"].shape[1]
        s.npair_cap = A["p_pairs"].shape[1]
        s.nl_cap = A["l_body"].shape[1]
        s.n_shapes = A["s_meta"].shape[0]
        s.nj_cap = A["j_def"].shape[1]
        for name in FIELD_ORDER:
            a = A[name]
            ptr = a.ctypes.data
            if not name.startswith("s_"):
                ptr += w * a.strides[0]
            setattr(s, name, ptr)
        self._s = s

    def step(self, dt, vel_iters, pos_iters, n_steps=1):
        rc = lib().orc_run(ctypes.byref(self._s), dt, self.inv_dt0, vel_iters, pos_iters, n_steps)
        if dt > 0.0 and n_steps > 0:
            self.inv_dt0 = 1.0 / dt
        return rc

    def collide_pair(self, sa, sb, xfa, xfb):
        xfa = np.ascontiguousarray(xfa, np.float64)
        xfb = np.ascontiguousarray(xfb, np.float64)
        mtype = np.zeros(1, np.int32)
        mf = np.zeros(4)
        pts = np.zeros(4)
        ids = np.zeros(2, np.uint32)
        lib().orc_collide_pair(ctypes.byref(self._s), sa, sb, xfa.ctypes.data, xfb.ctypes.data, mtype.ctypes.data,
                               mf.ctypes.data, pts.ctypes.data, ids.ctypes.data)
        return int(mtype[0]), mf, pts, ids

    def query_aabb(self, box, max_out=1024):
        """Fixture indices (ascending) whose fat AABB overlaps box = (lx, ly, ux, uy): World.query_aabb as a set."""
        box = np.ascontiguousarray(box, np.float64)
        out = np.zeros(max_out, np.int32)
        n = lib().orc_query_aabb(ctypes.byref(self._s), box.ctypes.data, out.ctypes.data, max_out)
        return out[:min(n, max_out)].copy()

    def ray_cast(self, p1, p2, max_out=1024):
        """(fixture indices, [fraction, nx, ny, px, py] rows) of every fixture the segment hits, ascending fixture index."""
        seg = np.ascontiguousarray([p1[0], p1[1], p2[0], p2[1]], np.float64)
        fix = np.zeros(max_out, np.int32)
        hit = np.zeros((max_out, 5))
        n = lib().orc_ray_cast(ctypes.byref(self._s), seg.ctypes.data, fix.ctypes.data, hit.ctypes.data, max_out)
        n = min(n, max_out)
        return fix[:n].copy(), hit[:n].copy()

    def test_point(self, f, p):
        pt = np.ascontiguousarray(p, np.float64)
        return bool(lib().orc_test_point(ctypes.byref(self._s), int(f), pt.ctypes.data))
