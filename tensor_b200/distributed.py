"""One-process-per-GPU plumbing for the batch-sharded hot path.

The batch axis is the only thing that is ever partitioned; ranks never exchange
tensor data.  The single exchange of the whole path is the scalar of
``dot_sum`` (a 1-element all-reduce: NCCL on GPUs, gloo in the CPU tests) plus
the max-over-ranks of the timing.  torch.distributed is plumbing only.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np

from ._lib import partition


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


_CPU_GROUP = None


def init_process_group(backend: str | None = None):
    import torch
    import torch.distributed as dist
    global _CPU_GROUP
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
            # a second, host-side group: a rank that waits in an NCCL barrier spins a kernel on its GPU, which
            # would disturb a rank that is timing work on ALL devices (bench.py's single-process section)
            _CPU_GROUP = dist.new_group(backend="gloo")
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def barrier_cpu():
    """Barrier that keeps every GPU idle while ranks wait (gloo over TCP on the host)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier(group=_CPU_GROUP) if _CPU_GROUP is not None else dist.barrier()


def strong_shard(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of `rank` when ONE global batch is split over `world` ranks —
    the same contiguous, 256-aligned partition the library uses across devices."""
    return partition(batch, world, rank)


def weak_first_item(per_rank: int, rank: int) -> int:
    """Global index of a rank's first item when every rank owns `per_rank` items."""
    return per_rank * rank


def all_reduce_scalar(value, np_dtype, device=None):
    """Sum one scalar over ranks.  Float/Double travel as float64, Int as wrapping int64."""
    import torch
    import torch.distributed as dist
    np_dtype = np.dtype(np_dtype)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np_dtype.type(value)
    if np_dtype == np.int64:
        t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    else:
        t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return np_dtype.type(t.item())


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
