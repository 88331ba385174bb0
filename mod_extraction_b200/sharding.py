"""Multi-GPU layout of the render path: one process per GPU, contiguous batch shards, no collective
while rendering (examples are independent: every delay line is zeroed per call, fx.py:92-93), one
all-gather at the end for the rendered batch and/or per-rank metrics."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_range(n_examples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split of the batch dimension; the first (n % world) ranks get one extra example."""
    assert 0 <= rank < world_size
    base, extra = divmod(n_examples, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_examples: int, world_size: int) -> List[int]:
    return [shard_range(n_examples, r, world_size)[1] - shard_range(n_examples, r, world_size)[0]
            for r in range(world_size)]


def all_gather_rendered(local: Tensor, n_examples: int, group=None) -> Tensor:
    """Gather the per-rank shards of a rendered tensor (first dim = examples) on every rank.
    NCCL over NVLink on GPUs (gloo in the CPU tests).  Ragged shards are padded to the largest."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_examples, world)
    assert local.size(0) == sizes[rank]
    mx = max(sizes)
    if local.size(0) < mx:
        pad = torch.zeros((mx - local.size(0),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], 0)


def max_over_ranks(value: float, device, group=None) -> float:
    """Device-timed durations are reported as the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def bind_to_gpu_numa_node(gpu_index: int) -> Optional[List[int]]:
    """One process per GPU: pin this process to the CPUs NVML reports as local to `gpu_index` BEFORE any pinned host
    buffer is allocated, so that the staging buffers of the host-buffer path land on the GPU's own NUMA node instead
    of wherever the launcher happened to start the process.  Returns the CPU list, or None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        cpus = [c for c in cpus if c < n_cpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
