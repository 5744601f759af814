import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "hostcheck")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(case):
    return np.load(os.path.join(GOLDEN, case + ".npz"))


def golden_scene_kw(g):
    """Scene keywords a fixture was generated with (oracle/gen_golden.py; absent = none)."""
    import json
    return json.loads(str(g["scene_kw"])) if "scene_kw" in g.files else {}


def arrays_from_golden(g, prefix="init_", base=None):
    """Schema arrays (fp64, batch of one) filled from a golden fixture's used rows."""
    from pypybox2d_b200 import schema
    caps = schema.Caps(*[int(x) for x in g["caps"]])
    if base is None:
        n_shapes = int(g["init_s_meta"].shape[0])
        A = schema.alloc(caps, 1, n_shapes, "f64")
        for f in ("s_meta", "s_radius", "s_data"):
            A[f][...] = g["init_" + f]
        src_prefix = "init_"
        for name, scope, tail, kind in schema.FIELDS:
            key = src_prefix + name
            if scope == "S" or key not in g.files:
                continue
            v = g[key]
            if scope == "W":
                A[name][0] = v
            else:
                A[name][0][:v.shape[0]] = v
    else:
        A = {k: v.copy() for k, v in base.items()}
    if prefix != "init_":
        for name, scope, tail, kind in schema.FIELDS:
            key = prefix + name
            if key not in g.files:
                continue
            v = g[key]
            if scope == "W":
                A[name][0] = v
            else:
                A[name][0][...] = 0
                A[name][0][:v.shape[0]] = v
    return A, caps


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()


_PARITY_REPORT = {}


@pytest.fixture(scope="session")
def parity_report():
    """Collects the measured parity errors of the -m gpu run; written to gpurun_out/parity_report.json (scratch) so that the
    tolerances pinned in the tests can be checked against what the hardware measured (a copy lives under profiles/)."""
    import json

    def add(key, value):
        _PARITY_REPORT[key] = value
    yield add
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(_PARITY_REPORT, f, indent=1, sort_keys=True, default=str)
    except Exception:
        pass
