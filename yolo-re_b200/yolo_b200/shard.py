"""Image-batch sharding across ranks.  Eval-mode inference has no cross-image
dependency (BN uses running stats, NMS is per image -- src/yolo/utils/nms.py:46), so the batch is
split contiguously, weights are replicated and the data path has NO collective; torch.distributed
is only used to agree on the timing (max over ranks) and to count images."""
from __future__ import annotations

import torch


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of the global batch owned by `rank` (first ranks take the remainder)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def aggregate_throughput(n_images_local: int, elapsed_s_local: float, device=None) -> tuple[float, int, float]:
    """(images/s of the whole job, total images, max elapsed seconds).  Uses all_reduce(SUM) on the
    image count and all_reduce(MAX) on the elapsed time when a process group is initialised."""
    import torch.distributed as dist
    n = torch.tensor([float(n_images_local)], dtype=torch.float64, device=device)
    t = torch.tensor([float(elapsed_s_local)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total, tmax = int(n.item()), float(t.item())
    return (total / tmax if tmax > 0 else 0.0), total, tmax


def bind_to_gpu_numa_node(device_index: int) -> list[int] | None:
    """One process per GPU: restrict this process to the CPU cores NVML reports as local to the GPU, so pinned host
    buffers allocated afterwards are first-touched on the GPU's NUMA node and the per-step H2D copy does not cross the
    socket interconnect.  Returns the core list, or None when NVML / affinity is unavailable (never raises)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        handle = None
        try:
            bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in range(ncpu) if (int(words[c // 64]) >> (c % 64)) & 1 and c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None
