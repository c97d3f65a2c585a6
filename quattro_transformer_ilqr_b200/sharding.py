"""Multi-GPU plumbing: contiguous trajectory shards, one process per GPU, one final gather.

Trajectories are independent MPC instances (no cross-trajectory arithmetic anywhere in the
reference, quattro_ilqr_tf/quattro_ilqr_tf.py:424-591), so the hot path has no collective: rank r
solves rows [lo, hi) of the global batch with its own ``BatchedILQR`` and the run ends with ONE
``all_gather`` of the per-trajectory results (costs, optionally iteration counts / alpha indices).
The same code runs over NCCL (GPU tensors) and gloo (CPU tensors, used by the tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous [lo, hi) of `total` items for `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(array, rank: int, world: int):
    lo, hi = shard_bounds(len(array), rank, world)
    return array[lo:hi]


def gather_rows(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather per-trajectory rows (dim 0) of every rank's shard into the global order.

    Equal shards use one ``all_gather_into_tensor`` (a single NCCL call); ragged shards are padded
    to the largest shard first.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = longest - local.shape[0]
    if pad:
        local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))])
    out = local.new_empty((world * longest,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if all(hi - lo == longest for lo, hi in sizes):
        return out
    return torch.cat([out[r * longest: r * longest + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def bind_host_to_gpu(device_index: int) -> bool:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA node / PCIe
    root), BEFORE pinned host buffers are allocated: pinned pages are first-touched by the allocating thread, so the
    host side of every H2D / D2H copy then lives in the memory of the socket the GPU hangs off.  With one process
    per GPU on a two-socket box this keeps half of the ranks from pushing their 40 GB/s of result traffic across the
    inter-socket link.  Best effort: returns False (and changes nothing) when NVML is not available."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = device_index
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            phys = int(vis.split(",")[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, bits in enumerate(mask) for b in range(64) if (int(bits) >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return False
        os.sched_setaffinity(0, allowed)
        return True
    except Exception:
        return False
