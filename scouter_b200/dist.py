"""Data-parallel plumbing for the forward path (reference: train.py:139-160, tools/prepare_things.py:9-31).

SlotModel.forward shards over the batch with no collective: every rank holds a full replica and runs its own images.
The only cross-rank operations are the rendezvous, a barrier, and reductions of *measurements* (max time over ranks,
sum of images).  `torch.distributed` with NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None, device: torch.device | None = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's env; initialises the process group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend in (None, "nccl") and device is not None and device.type == "cuda":
            kw["device_id"] = device
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of `total` independent images for `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_images_per_second(images_this_rank: int, ms_this_rank: float, device: torch.device | str = "cpu") -> float:
    """Whole-job throughput: all images of all ranks / the slowest rank's time."""
    return sum_over_ranks(images_this_rank, device) / (max_over_ranks(ms_this_rank, device) / 1e3)
