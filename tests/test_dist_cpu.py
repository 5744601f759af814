"""Multi-GPU host logic on CPU: world sharding and the end-of-run gather over gloo (world_size 2)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from pypybox2d_b200 import dist, scenes


def test_shard_ranges_partition_the_world_index():
    for n in (1, 7, 8, 1000, 1048576):
        for ws in (1, 2, 3, 8):
            r = [dist.shard_range(n, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(ws - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_jitter_depends_only_on_the_global_world_index():
    import torch
    idx = np.arange(2 * 211, dtype=np.uint64)
    a = scenes.u01(3, 123457, idx)
    w = torch.tensor([[[123457]]], dtype=torch.int64)
    b = scenes.u01_torch(3, w, torch.arange(2 * 211, dtype=torch.int64).view(1, 211, 2)).numpy().reshape(-1)
    assert np.array_equal(a, b)
    assert 0.0 <= a.min() and a.max() < 1.0 and abs(a.mean() - 0.5) < 0.05
    assert np.array_equal(scenes.jitter(3, 0, 5), np.zeros((5, 2)))


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import torch
from pypybox2d_b200 import dist
rank, local_rank, ws = dist.init(backend="gloo")
n = 11
a, b = dist.shard_range(n, rank, ws)
rows = torch.arange(a, b, dtype=torch.float64).view(-1, 1) * torch.tensor([[1.0, 10.0, 100.0, 1000.0]], dtype=torch.float64)
full = dist.gather_rows(rows, n, rank, ws)
assert full.shape == (n, 4), full.shape
assert torch.equal(full[:, 0], torch.arange(n, dtype=torch.float64)), full[:, 0]
t = dist.max_over_ranks(1.0 + rank, torch.device("cpu"))
assert t == float(ws), t
dist.barrier()
print("rank %%d ok" %% rank)
"""


def test_gather_rows_over_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
