"""Multi-GPU: contiguous batch shards, one process per GPU, no collective on the hot path.

Jacobians of different samples are independent and the plan is tiny, so every
rank uploads its own copy of the plan and evaluates ``batch_per_rank`` samples
(weak scaling) or its slice of a fixed batch.  ``gather_for_verification`` is the
only collective: an all-gather of result shards (NCCL over NVLink on the GPU box,
gloo in the CPU tests), used to assemble results for checking - never in a timed
region.
"""
from __future__ import annotations

import os
from typing import List, Tuple


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of this rank's contiguous shard; shards differ by at most one sample."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def env_rank() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process per GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: str = "nccl"):
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            import torch
            kwargs["device_id"] = torch.device(f"cuda:{local_rank}")   # no "guessing device ID" at the first barrier
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def gather_for_verification(local, sizes: List[int]):
    """All-gather variable-sized shards along dim 0 (verification only)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    longest = max(sizes)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(parts, pad)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
