"""Multi-GPU plumbing: worlds are independent, so the batch is sharded contiguously over
ranks with NO per-step collective.  The only exchange is one
all-reduce(sum) per rollout over a tiny episode-statistics vector -- NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard(total_worlds: int, rank: int, world_size: int) -> tuple[int, int]:
    """(offset, count) of the contiguous block of worlds owned by ``rank``.  World b always
    draws from Philox(seed, world=b) so results do not depend on the number of GPUs."""
    if total_worlds < 0 or not (0 <= rank < world_size):
        raise ValueError("bad shard request")
    base, rem = divmod(total_worlds, world_size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Join the process group torchrun described (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)  # binds the communicator to this rank's GPU
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Once-per-rollout reduction of the episode statistics vector (sum over ranks)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule: a multi-GPU duration is the max over ranks."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
