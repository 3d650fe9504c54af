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


def bind_to_local_numa_node(device: int) -> str:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE pinned host buffers are allocated:
    page-locked memory is then node-local (first touch), and host<->device copies of the ranks of one box stop
    crossing the socket interconnect.  Returns a one-line description; a no-op when the topology is not exposed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                 # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown (single node or not exposed)"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return f"numa node {node}: none of its CPUs is available to this process"
        os.sched_setaffinity(0, allowed)
        return f"numa node {node}, {len(allowed)} CPUs"
    except Exception as e:      # noqa: BLE001  (topology files missing in a container, no NVML, ...)
        return f"not bound ({type(e).__name__})"


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
