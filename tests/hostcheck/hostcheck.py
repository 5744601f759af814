"""Loader for the TEST-ONLY host instantiation of the device step (tests/hostcheck/hostcheck.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))

FIELD_ORDER = ("w_count,w_flags,w_status,w_gravity,w_island,"
               "b_xf,b_pos,b_vel,b_pos0,b_force,b_mass,b_damp,b_flags,b_island,"
               "f_body,f_shape,f_mat,f_filter,f_fat,f_tight,f_moved,"
               "c_fix,c_flags,c_mat,c_mtype,c_mf,c_pts,c_imp,c_id,"
               "p_pairs,l_body,l_contact,l_start,s_meta,s_radius,s_data,"
               "w_joint,j_def,j_par,j_imp,j_tmp,j_adj,b_jadj,l_joint,l_jstart").split(",")


class Caps(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n_worlds", "nb", "nf", "nc", "npair", "nl", "n_shapes", "real_bytes",
                                              "device", "threads", "ctas_per_sm", "reserved", "nj", "joint_types")]


class Buffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in FIELD_ORDER]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libhostcheck.so")
        srcs = [os.path.join(_HERE, "hostcheck.cpp")] + [
            os.path.join(_ROOT, "pypybox2d_b200", "csrc", f) for f in ("b2b_step.cuh", "b2b_collide.cuh", "b2b_math.cuh", "b2b_joints.inl", "b2b_toi.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-builtin-pow", "-o", so,
                                   srcs[0], "-lm"])
        _LIB = ctypes.CDLL(so)
        _LIB.hc_step.argtypes = [ctypes.POINTER(Caps), ctypes.POINTER(Buffers), ctypes.c_double, ctypes.c_double,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _LIB.hc_step.restype = ctypes.c_int
        _LIB.hc_joint_world_steps.argtypes = [ctypes.c_int]
        _LIB.hc_joint_world_steps.restype = ctypes.c_longlong
        _LIB.hc_replans.argtypes = [ctypes.c_int]
        _LIB.hc_replans.restype = ctypes.c_longlong
    return _LIB


def step(arrays, dt, inv_dt0, vel_iters, pos_iters, n_steps=1):
    """Step every world of `arrays` (host numpy, f32 or f64 reals) in place with the host-compiled device source."""
    A = arrays
    c = Caps()
    c.n_worlds = A["w_flags"].shape[0]
    c.nb = A["b_flags"].shape[1]
    c.nf = A["f_body"].shape[1]
    c.nc = A["c_flags"].shape[1]
    c.npair = A["p_pairs"].shape[1]
    c.nl = A["l_body"].shape[1]
    c.n_shapes = A["s_meta"].shape[0]
    c.real_bytes = A["b_xf"].dtype.itemsize
    c.nj = A["j_def"].shape[1] if int(A["w_joint"][:, 0].max()) > 0 else 0
    c.joint_types = 0
    for w in range(A["w_joint"].shape[0]):
        for t in A["j_def"][w, :int(A["w_joint"][w, 0]), 0]:
            c.joint_types |= 1 << int(t)
    b = Buffers()
    for n in FIELD_ORDER:
        assert A[n].flags["C_CONTIGUOUS"], n
        setattr(b, n, A[n].ctypes.data)
    return lib().hc_step(ctypes.byref(c), ctypes.byref(b), dt, inv_dt0, vel_iters, pos_iters, n_steps)


def joint_world_steps(reset=True):
    """World-steps so far whose plan row carried a joint island schedule (phase_schedule in b2b_step.cuh)."""
    return int(lib().hc_joint_world_steps(1 if reset else 0))


def replans(reset=True):
    """World-steps so far that entered the step with invalid island lists (each one rebuilds lists and schedule)."""
    return int(lib().hc_replans(1 if reset else 0))
