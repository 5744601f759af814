"""Multi-GPU plumbing: one process per GPU, worlds sharded by contiguous index blocks.

Worlds never interact (each reference `World` owns its own `ContactManager`, world.py:82), so the
step path has NO collective.  The only exchange is the optional end-of-run gather of per-world
checksums / energies, done with `torch.distributed.all_gather_into_tensor` -- NCCL over
NVLink on GPUs, gloo on CPU for the tests.
"""
import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n_worlds, rank, world_size):
    """Contiguous block [start, stop) of the global world index owned by `rank` (sizes differ by at most 1)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_worlds), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world_size = env_rank()
    if world_size == 1:
        return rank, local_rank, world_size
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world_size, **kw)
    return rank, local_rank, world_size


def gather_rows(local_rows, n_global, rank, world_size):
    """All-gather per-world rows ([n_local, k]) into [n_global, k] in global world order on every rank.
    Shards may differ in size by one row, so they are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return local_rows
    k = local_rows.shape[1]
    sizes = [shard_range(n_global, r, world_size) for r in range(world_size)]
    pad = max(b - a for a, b in sizes)
    buf = torch.zeros((pad, k), dtype=local_rows.dtype, device=local_rows.device)
    buf[:local_rows.shape[0]] = local_rows
    out = torch.empty((world_size * pad, k), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, buf)
    return torch.cat([out[r * pad:r * pad + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing: max over ranks, never wall clock)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
