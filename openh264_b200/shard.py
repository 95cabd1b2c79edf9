"""Multi-GPU plumbing for batched independent streams: replicas only, no data-path collective.

One process per GPU (torchrun).  A job of `total` streams is dealt round-robin — stream s belongs to rank
s mod world — and the only communication is the timing contract: a barrier before/after the timed region and
MAX over ranks of the per-rank step time (device-measured), SUM of the pictures coded."""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def streams_for_rank(total, rank, world):
    """ids of the streams rank `rank` codes: s mod world == rank (every stream exactly once, sizes differ by <= 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, total, world))


def init(backend, device=None):
    rank, local, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def barrier():
    if dist.is_initialized():
        dist.barrier()


def job_totals(seconds, pictures, device="cpu"):
    """(max over ranks of `seconds`, sum over ranks of `pictures`): the whole-job figures bench.py reports."""
    if not dist.is_initialized():
        return float(seconds), int(pictures)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    n = torch.tensor([int(pictures)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())
