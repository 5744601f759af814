"""ctypes binding of libb2b.so (include/b2b.h).  There is no fallback: if the CUDA library is
missing or a call fails, this module raises."""
import ctypes
import os
import subprocess

from .schema import FIELD_NAMES

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("B2B_LIB") or os.path.join(CSRC, "libb2b.so")      # (B2B_LIB: an experimental build, tools/build_variant.sh)


class Caps(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n_worlds", "nb", "nf", "nc", "npair", "nl", "n_shapes", "real_bytes",
                                              "device", "threads", "ctas_per_sm", "reserved", "nj", "joint_types")]


class Buffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in FIELD_NAMES]


class B2BError(RuntimeError):
    pass


_LIB = None


def build(verbose=False):
    """Compile csrc/ for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "libb2b.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise B2BError("building libb2b.so failed")
    return SO_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise B2BError("CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback for the step)" % SO_PATH)
    L = ctypes.CDLL(SO_PATH)
    vp, i32, u32, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_double
    L.b2b_create.argtypes = [ctypes.POINTER(Caps), ctypes.POINTER(vp)]
    L.b2b_destroy.argtypes = [vp]
    L.b2b_destroy.restype = None
    L.b2b_scratch_bytes.argtypes = [ctypes.POINTER(Caps)]
    L.b2b_scratch_bytes.restype = ctypes.c_size_t
    L.b2b_bind.argtypes = [vp, ctypes.POINTER(Buffers)]
    L.b2b_init_proxies.argtypes = [vp, vp]
    L.b2b_step.argtypes = [vp, dbl, dbl, i32, i32, i32, u32, vp]
    L.b2b_step_host.argtypes = [vp, vp, vp, dbl, dbl, i32, i32, i32, u32, vp]
    L.b2b_solve_islands.argtypes = [vp, dbl, dbl, i32, i32, u32, vp]
    L.b2b_collide_pairs.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.b2b_query_aabb.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp]
    L.b2b_ray_cast.argtypes = [vp, i32, vp, vp, i32, i32, vp, vp, vp, vp]
    L.b2b_test_points.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.b2b_status.argtypes = [vp, ctypes.POINTER(u32), vp]
    L.b2b_checksum.argtypes = [vp, vp, vp]
    L.b2b_profile.argtypes = [vp, ctypes.POINTER(ctypes.c_int64), i32]
    L.b2b_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)]
    L.b2b_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                           ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64)]
    for f in ("b2b_last_error", "b2b_field_names", "b2b_version"):
        getattr(L, f).restype = ctypes.c_char_p
    _LIB = L
    return L


def check(rc, what, allow_overflow=False):
    if rc == 0:
        return
    if allow_overflow and rc == -4:
        return
    raise B2BError("%s failed (%d): %s" % (what, rc, lib().b2b_last_error().decode()))
