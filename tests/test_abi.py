"""The C-ABI library loads and exports every symbol include/b2b.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from pypybox2d_b200 import _lib, schema


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b2b.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2b_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("b2b_create", "b2b_destroy", "b2b_bind", "b2b_step", "b2b_step_host", "b2b_init_proxies",
                 "b2b_collide_pairs", "b2b_solve_islands", "b2b_status", "b2b_checksum", "b2b_kernel_ms", "b2b_last_error", "b2b_scratch_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in declared_functions():
        assert hasattr(L, name), "libb2b.so does not export %s" % name


def test_field_order_is_shared_by_header_library_and_schema():
    L = _lib.lib()
    assert L.b2b_field_names().decode().split(",") == schema.FIELD_NAMES
    src = open(os.path.join(ROOT, "include", "b2b.h")).read()
    body = src[src.index("typedef struct b2b_buffers"):src.index("} b2b_buffers;")]
    members = re.findall(r"\*([a-z_0-9]+)", body)
    assert members == schema.FIELD_NAMES
    assert ctypes.sizeof(_lib.Buffers) == 8 * len(schema.FIELD_NAMES)


def test_scratch_bytes_and_bad_caps():
    L = _lib.lib()
    c = _lib.Caps()
    assert L.b2b_scratch_bytes(ctypes.byref(c)) == 0          # all-zero caps are rejected
    c.n_worlds, c.nb, c.nf, c.nc, c.npair, c.nl, c.n_shapes, c.real_bytes = 4, 211, 211, 908, 1816, 1119, 2, 4
    n32 = L.b2b_scratch_bytes(ctypes.byref(c))
    c.real_bytes = 8
    n64 = L.b2b_scratch_bytes(ctypes.byref(c))
    assert 0 < n32 < n64
    ctx = ctypes.c_void_p()
    c.real_bytes = 3
    assert L.b2b_create(ctypes.byref(c), ctypes.byref(ctx)) == -1
    assert b"bad capacities" in L.b2b_last_error()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pypybox2d_b200 as b2
    w = b2.scenes.hello(b2)
    with pytest.raises(_lib.B2BError):
        w.step(0.1, 10, 10)
    L = _lib.lib()
    c = _lib.Caps()
    c.n_worlds, c.nb, c.nf, c.nc, c.npair, c.nl, c.n_shapes, c.real_bytes = 1, 2, 2, 8, 16, 10, 2, 4
    ctx = ctypes.c_void_p()
    assert L.b2b_create(ctypes.byref(c), ctypes.byref(ctx)) == -2     # B2B_ERR_CUDA
    assert ctx.value is None


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pypybox2d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "b2_oracle" not in text and "import oracle" not in text and "libb2oracle" not in text, f
