"""Job sharding across the GPUs of one box.

The reference's only parallelism is one Celery task per job
(/root/reference/backend/app/workers/celery_app.py:7-21): jobs never exchange data.  The same holds
here: clips are partitioned across ranks (longest-processing-time greedy, which degenerates to
round-robin for equal lengths), every rank runs its own plan on its own GPU, and NO collective
touches the data path.  ``reduce_stats`` is an optional all-reduce of three scalars for reporting.
"""
from __future__ import annotations

from typing import List, Sequence


def partition(lengths: Sequence[int], n_shards: int) -> List[List[int]]:
    """Indices of the clips each shard processes; deterministic, balanced by total samples."""
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * n_shards
    shards: List[List[int]] = [[] for _ in range(n_shards)]
    for i in order:
        s = min(range(n_shards), key=lambda k: (loads[k], k))
        shards[s].append(i)
        loads[s] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards


def local_shard(lengths: Sequence[int], rank: int, world_size: int) -> List[int]:
    return partition(lengths, world_size)[rank]


def reduce_stats(audio_seconds: float, elapsed_s: float, n_bytes: float, group=None):
    """(sum audio-seconds, max elapsed, sum bytes) over ranks; works on gloo (CPU) and nccl."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return audio_seconds, elapsed_s, n_bytes
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    sums = torch.tensor([audio_seconds, n_bytes], dtype=torch.float64, device=dev)
    mx = torch.tensor([elapsed_s], dtype=torch.float64, device=dev)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return float(sums[0]), float(mx[0]), float(sums[1])


def bind_to_gpu_numa(device_index: int) -> List[int]:
    """Pin this process to the CPUs closest to `device_index` (NVML's CPU affinity of the GPU).

    One process per GPU stages its clips through pinned host buffers; Linux places those pages on the
    NUMA node of the CPU that first touches them, so a rank running on the far socket pushes every
    byte over the inter-socket link as well as PCIe.  Call this BEFORE allocating pinned memory.
    Returns the CPU list it bound to ([] when NVML or the affinity call is unavailable -- binding is an
    optimisation, never a requirement)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
            n_words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        finally:
            pynvml.nvmlShutdown()
        cpus = [w * 64 + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []
